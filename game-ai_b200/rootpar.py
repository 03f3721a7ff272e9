"""Root/leaf parallelism across GPUs: sharding arithmetic and the once-per-move merge.

The path shards with NO data-path collective: every rank plays its own leaves (bench) or its own tree
(player); the only exchange is a sum of the root's per-move statistics -- `visits` u32 and `value[2]` f32 per root move,
<= 96 moves (MoveNode::numVisited / value[], MCTSPlayer.h:47-69) -- once per move decision, after which every rank applies
the reference's final-move rule (selectBestMove, MCTSPlayer.cpp:601-632) to identical numbers with the same tie-break
seed and therefore picks the same move.  On GPUs the sum is `gai_allreduce_root` (NCCL, C ABI); this module is the same
logic over torch.distributed so that it is testable with the gloo backend on CPU (tests/test_rootpar_gloo.py).
"""
import numpy as np

MAX_ROOT_MOVES = 96


def shard_plan(rank, world, leaves_per_rank, playouts_per_leaf, base_key):
    """Weak scaling: every rank owns `leaves_per_rank` leaves.  Playout ids are globally unique and contiguous per rank,
    the corpus key differs per rank (independent synthetic shards)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    per_rank = leaves_per_rank * playouts_per_leaf
    return {"first_playout_id": rank * per_rank, "corpus_key": base_key + rank, "playouts": per_rank,
            "global_playouts": per_rank * world}


def merge_root_stats(visits, values, dist=None, group=None):
    """Sum root-child statistics over all ranks (in a fixed 96-slot buffer, like the C ABI).  -> (visits, values)"""
    import torch
    visits = np.asarray(visits, dtype=np.uint32); values = np.asarray(values, dtype=np.float32)
    n = visits.shape[0]
    if n > MAX_ROOT_MOVES or values.shape != (n, 2):
        raise ValueError("root statistics: at most 96 moves, values of shape (n, 2)")
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return visits.copy(), values.copy()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    v = torch.zeros(MAX_ROOT_MOVES, dtype=torch.int64, device=dev); v[:n] = torch.from_numpy(visits.astype(np.int64))
    f = torch.zeros(MAX_ROOT_MOVES, 2, dtype=torch.float32, device=dev); f[:n] = torch.from_numpy(values)
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group)
    return v[:n].cpu().numpy().astype(np.uint32), f[:n].cpu().numpy()


def select_best_move(visits, values, player, eps, seed):
    """The reference's final-move rule: best value/(1+n), uniformly among those within eps (MCTSPlayer.cpp:601-618)."""
    best, bestv = [], 0.0
    for m in range(len(visits)):
        v = np.float32(values[m][player]) / np.float32(1.0 + visits[m])
        if not best:
            best, bestv = [m], v
            continue
        diff = v - bestv
        if abs(diff) < eps:
            best.append(m)
        elif diff > 0:
            best, bestv = [m], v
    if not best:
        return 0
    rng = np.random.RandomState(seed)
    return best[rng.randint(len(best))] if len(best) > 1 else best[0]
