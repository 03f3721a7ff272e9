// game-ai_b200/host/gpu_mcts_player.cpp -- GPU-backed MCTS player plugin, exports createPlayer.
//
// Drop-in variant of the reference's MC::Player (c++/MCTSPlayer/MCTSPlayer.{h,cpp}, factory MCplayer_factory.cpp:46-82):
// same IGamePlayer vtable, same config keys, same tree policy --
//   edge value  = value[player] * probability / (1 + visits)                      (MoveNode::getWeight, MCTSPlayer.h:55-58)
//   UCB         = value + C * sqrt(ln(N + 1) / (1 + visits)), ties broken uniformly (Player::selectMove, MCTSPlayer.cpp:564-599)
//   back-up     = value added to every edge of the path, scaled by (1 - wb) + wb * probability on the way up (:496-507)
//   final move  = best VALUE, uniformly among those within best_move_value_eps       (selectBestMove, :601-632)
//   states reached by different move orders share one node                          (makeTreeNode, :634-662)
// but the simulation phase (the part of selection_playOut, :387-470, below the tree) runs in libgameai_gpu.so
// (gameai_gpu.h): selection / expansion / back-propagation stay here on the host.  Three ways to feed the GPU:
//
//   expand_all=1 (default)  batch-by-expansion.  A selection ends at the first node whose moves have no statistics yet;
//       ALL its successors are played out `playouts_per_leaf` times each in one device pass
//       (gai_loa_rollout_children_submit: move generation + apply on the device, results in move order).  This is the
//       batched form of what UCB does anyway -- every move of a fresh node is tried before any is tried twice
//       (:582-593) -- and it is what takes the host off the critical path: one selection (3 us of host time) now buys
//       ~30 x playouts_per_leaf playouts instead of playouts_per_leaf.
//   expand_all=0            one new leaf per selection, `playouts_per_leaf` playouts from it (round-1 behaviour).
//   expand_all=0 playouts_per_leaf=1   the reference-like mode used for same-budget comparisons: one playout per
//       simulation, and `expand_size` is honoured exactly -- the first expand_size positions of the playout below the
//       leaf become permanent nodes too (expansion, :524-562); the host replays them from the Philox choice stream.
//
// Batches are pipelined (`pipeline_depth` in flight, ticketed C-ABI calls): while batch k plays out, batch k+1 is
// selected -- the leaves of k already carry their virtual loss (their playouts are charged as visits before the
// values arrive), so k+1 explores elsewhere.  Other deliberate differences from the reference: visit counters are
// 32-bit (the reference's unsigned char wraps at 255, MCTSPlayer.h:51); the ply cap counts from the leaf, not from
// the root (:445); draws / capped playouts back up 0.5 like the reference's Score / constant eval
// (LinesOfActionZasady.cpp:382-397,420-432).
//
// Extra config keys: gpu_batch (selections per batch), playouts_per_leaf, expand_all, pipeline_depth, devices ("0" or
// "0,1,2,3": root parallelism, one independent tree per GPU merged at the root), trees_per_gpu (that many trees on every
// listed GPU), philox_seed.  Multi-process root
// parallelism (one rank per GPU): WORLD_SIZE / RANK / GAI_NCCL_ID_FILE in the environment -> gai_allreduce_root over
// NCCL once per move decision.  Unknown keys are an error; keys starting with '_' are the reference's way of
// commenting an attribute out (run_config_loa.xml: _move_sim_limit) and are skipped.
#include "gameai/GamePlayer.h"
#include "gameai/GameRules.h"
#include "gameai/philox.h"
#include "gameai_gpu.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

using Clock = std::chrono::steady_clock;

struct Edge { int32_t child = -1; uint32_t visits = 0; float value[2] = { 0.f, 0.f }; float prob = 1.f; };
struct Node {
	GameState* state = nullptr; MoveList* moves = nullptr;
	int first_edge = 0, n_edges = -1;        // n_edges < 0: moves not generated yet
	uint32_t visits = 0, visit_id = 0;
	uint8_t cp = 0; bool terminal = false;
	bool pending = false;                    // expand_all: its children are being played out right now
	bool expanded = false;                   // expand_all: its edges carry statistics
	int score[2] = { 0, 0 };
};
struct Key { uint64_t w, b, cp; bool operator==(const Key& o) const { return w == o.w && b == o.b && cp == o.cp; } };
struct KeyHash { size_t operator()(const Key& k) const { uint64_t h = k.w * 0x9E3779B97F4A7C15ull ^ (k.b + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull ^ k.cp; h ^= h >> 29; return (size_t)h; } };

// Transposition table (makeTreeNode, MCTSPlayer.cpp:634-662, keyed there by the ToString text): open addressing, one
// 32-byte entry per position, no allocation per insert.
struct FlatTable {
	struct Entry { Key k; int32_t idx; uint32_t used; };
	std::vector<Entry> slots; size_t mask = 0, count = 0;
	void init(size_t cap) { size_t c = 1; while (c < cap) c <<= 1; slots.assign(c, Entry{ Key{ 0, 0, 0 }, -1, 0 }); mask = c - 1; count = 0; }
	void clear() { if (!slots.empty()) init(slots.size()); }
	int find(const Key& k) const
	{
		if (slots.empty()) return -1;
		for (size_t i = KeyHash()(k) & mask;; i = (i + 1) & mask) {
			const Entry& e = slots[i];
			if (!e.used) return -1;
			if (e.k == k) return e.idx;
		}
	}
	void emplace(const Key& k, int idx)
	{
		if (slots.empty()) init(1u << 19);
		if (2 * (count + 1) > slots.size()) {                                // keep the load factor under 1/2
			std::vector<Entry> old; old.swap(slots);
			init(old.size() * 2);
			for (const Entry& e : old) if (e.used) emplace(e.k, e.idx);
		}
		for (size_t i = KeyHash()(k) & mask;; i = (i + 1) & mask) {
			Entry& e = slots[i];
			if (!e.used) { e.k = k; e.idx = idx; e.used = 1; ++count; return; }
			if (e.k == k) { e.idx = idx; return; }
		}
	}
};

struct Settings {
	int player = 0, num_players = 2;
	float C = 1.0f, eps = 0.005f;
	int max_plies = 20;
	int batch = 1024, ppl = 32;
	bool expand_all = true; int pipeline = 2;
	int expand_size = 1;                      // NodesToAppendDuringExpansion (MCplayer_factory.cpp:54)
	float cycle_value = 0.5f;                 // cycle_score / 100 (MCTSConfig::CycleScore, :62)
	float weighted_backprop = 0.f;            // WeightedBackprop (:60)
	std::string eval_function, trace, out_dir;
	double time_limit = 1.0; long sim_limit = -1;
	uint64_t philox_key = 0x10A5EED;
	unsigned seed = 1;
	std::string name;
};

// One independent search tree bound to one GPU context (root parallelism = several of these).
struct Tree {
	IGameRules* rules = nullptr;
	gai_ctx* ctx = nullptr;
	Settings cfg;
	std::vector<Node> nodes;
	std::vector<Edge> edges;
	FlatTable table;
	std::mt19937 rng;
	uint32_t visit_id = 0;
	uint64_t next_playout_id = 0, key = 0;
	// per-search statistics
	uint64_t playouts = 0, plies = 0, batches = 0, capped = 0, selections = 0, collisions = 0;
	double gpu_ms = 0;
	Histogram<long>* path_size = nullptr;
	std::string error;

	Key key_of(const GameState* s) const
	{
		const uint32_t* h = rules->GetStateHash(s);          // LOA: the 24-byte {white, black, current_player}
		Key k; memcpy(&k, h, 24); k.cp &= 1; return k;
	}
	int make_node(GameState* s)                               // takes ownership of s
	{
		const Key k = key_of(s);
		const int known = table.find(k);
		if (known >= 0) { rules->ReleaseGameState(s); return known; }
		if (nodes.capacity() == 0) { nodes.reserve(1u << 19); edges.reserve(1u << 23); }      // no regrowth mid-search (untouched pages cost nothing)
		Node n; n.state = s; n.cp = (uint8_t)rules->GetCurrentPlayer(s);
		n.terminal = rules->IsTerminal(s);
		if (n.terminal) { rules->Score(s, n.score); n.n_edges = 0; }
		nodes.push_back(n);
		const int idx = (int)nodes.size() - 1;
		table.emplace(k, idx);
		return idx;
	}
	// Move generation is deferred to the first descent INTO a node: most leaves of the one-leaf-per-selection modes are
	// never selected again.
	void expand(Node& n)
	{
		if (n.n_edges >= 0) return;
		n.moves = rules->GetPlayerLegalMoves(n.state, n.cp);
		n.n_edges = rules->GetNumMoves(n.moves);
		n.first_edge = (int)edges.size();
		edges.resize(edges.size() + (size_t)n.n_edges);
		if (cfg.num_players > 2 || cfg.weighted_backprop != 0.f)       // LOA: every move has probability 1 (no lookup needed)
			for (int i = 0; i < n.n_edges; ++i) edges[n.first_edge + i].prob = std::get<1>(rules->GetMoveFromList(n.moves, i));
	}
	void clear()
	{
		for (auto& n : nodes) { if (n.moves) rules->ReleaseMoveList(n.moves); rules->ReleaseGameState(n.state); }
		nodes.clear(); edges.clear(); table.clear();
	}
	// Tree reuse with garbage collection (reference: findRootNode + freeSubTree, MCTSPlayer.cpp:219-251,330-354):
	// keep what is reachable from the new root, free the rest, compact the arenas.  Returns the new root index.
	int collect(int root)
	{
		std::vector<int> remap(nodes.size(), -1), order;
		order.reserve(nodes.size() / 4 + 16);
		remap[root] = 0; order.push_back(root);
		for (size_t h = 0; h < order.size(); ++h) {                       // breadth first over the DAG
			const Node& n = nodes[order[h]];
			for (int i = 0; i < n.n_edges; ++i) {
				const int c = edges[n.first_edge + i].child;
				if (c >= 0 && remap[c] < 0) { remap[c] = (int)order.size(); order.push_back(c); }
			}
		}
		std::vector<Node> nn; nn.reserve(std::max(order.size() * 2, (size_t)1 << 19));
		std::vector<Edge> ne; ne.reserve(std::max(edges.size() / 2 + 64, (size_t)1 << 23));
		for (int old : order) {
			Node n = nodes[old];
			const int fe = (int)ne.size();
			for (int i = 0; i < n.n_edges; ++i) { Edge e = edges[n.first_edge + i]; if (e.child >= 0) e.child = remap[e.child]; ne.push_back(e); }
			n.first_edge = fe; n.visit_id = 0;
			nn.push_back(n);
		}
		for (size_t i = 0; i < nodes.size(); ++i) if (remap[i] < 0) { if (nodes[i].moves) rules->ReleaseMoveList(nodes[i].moves); rules->ReleaseGameState(nodes[i].state); }
		nodes.swap(nn); edges.swap(ne);
		table.clear(); visit_id = 0;
		for (size_t i = 0; i < nodes.size(); ++i) table.emplace(key_of(nodes[i].state), (int)i);
		return 0;
	}

	// UCB choice among the moves that do not lead back into the current path (MCTSPlayer.cpp:564-599).  The values live in
	// the node's contiguous edge array; whether a child lies on the current path is checked only for the edge that WON
	// (then that edge is excluded and the choice repeated).
	int choose_edge(const Node& n)
	{
		const float lnN = (float)std::log((double)n.visits + 1.0);
		const int ne = n.n_edges < 128 ? n.n_edges : 128;                 // <= 96 moves (LinesOfActionZasady.cpp:226)
		float val[128];
		const Edge* ed = edges.data() + n.first_edge;
		const int cp = n.cp;
		for (int i = 0; i < ne; ++i) {                                    // no branches: the compiler vectorises sqrt and divide (8 lanes)
			const float oo = 1.0f / (1.0f + (float)ed[i].visits);
			val[i] = ed[i].value[cp] * ed[i].prob * oo + cfg.C * std::sqrt(lnN * oo);
		}
		for (;;) {
			float best = -1e30f; int ties[128]; int nt = 0;
			for (int i = 0; i < ne; ++i) {
				const float v = val[i];
				if (v > best) { best = v; nt = 0; ties[nt++] = i; }
				else if (v == best) ties[nt++] = i;
			}
			if (nt == 0 || best == -1e30f) return -1;
			const int pick = ties[nt == 1 ? 0 : std::uniform_int_distribution<int>(0, nt - 1)(rng)];
			const int child = ed[pick].child;
			if (child < 0 || nodes[child].visit_id != visit_id) return pick;
			val[pick] = -1e30f;                                           // would close a loop: choose again without it
		}
	}

	struct Pending {
		std::vector<int> path;       // edge indices from the root
		std::vector<int> npath;      // node indices from the root (the node each edge leaves from), then the leaf
		int leaf = -1;
	};

	// value added to every edge of the path, scaled on the way up (MCTSPlayer.cpp:496-507)
	void backprop(const std::vector<int>& path, float vw, float vb)
	{
		const float v1 = 1.0f - cfg.weighted_backprop, v2 = 1.0f - v1;
		for (size_t i = path.size(); i-- > 0;) {
			Edge& e = edges[path[i]];
			e.value[0] += vw; e.value[1] += vb;
			if (v2 != 0.f) { const float f = v1 + v2 * e.prob; vw *= f; vb *= f; }
		}
	}
	void charge(const Pending& p, uint32_t visits)
	{
		for (int ei : p.path) edges[ei].visits += visits;
		for (int ni : p.npath) nodes[ni].visits += visits;
	}
	enum Sel { LEAF, RESOLVED, COLLISION };
	// One selection.  LEAF: p.leaf must be played out (its playouts are already charged as virtual visits along the
	// path); RESOLVED: the simulation ended on the host (terminal node, cycle, dead end) and was backed up;
	// COLLISION (expand_all): the descent ran into a node whose expansion is still in flight -- nothing was charged.
	Sel select_leaf(int root, Pending& p, long& sims)
	{
		++visit_id; ++selections;
		p.path.clear(); p.npath.clear(); p.leaf = -1;
		int ni = root;
		for (;;) {
			Node& n = nodes[ni];
			n.visit_id = visit_id;
			p.npath.push_back(ni);
			if (n.terminal) {
				charge(p, (uint32_t)cfg.ppl); sims += cfg.ppl;
				backprop(p.path, n.score[0] / 100.0f * cfg.ppl, n.score[1] / 100.0f * cfg.ppl);
				return RESOLVED;
			}
			if (cfg.expand_all) {
				if (n.pending) { ++collisions; return COLLISION; }
				if (!n.expanded) {
					expand(n);
					if (n.n_edges == 0) { charge(p, (uint32_t)cfg.ppl); sims += cfg.ppl; backprop(p.path, cfg.cycle_value * cfg.ppl, cfg.cycle_value * cfg.ppl); return RESOLVED; }
					n.pending = true;
					p.leaf = ni;
					charge(p, (uint32_t)(n.n_edges * cfg.ppl)); sims += (long)n.n_edges * cfg.ppl;
					return LEAF;
				}
			} else expand(n);
			const int ci = choose_edge(n);
			if (ci < 0) {                                                  // every move closes a loop: CycleScore (MCTSPlayer.cpp:489-491)
				charge(p, (uint32_t)cfg.ppl); sims += cfg.ppl;
				backprop(p.path, cfg.cycle_value * cfg.ppl, cfg.cycle_value * cfg.ppl);
				return RESOLVED;
			}
			const int ei = n.first_edge + ci;
			p.path.push_back(ei);
			if (edges[ei].child >= 0) { ni = edges[ei].child; continue; }
			auto [mv, prob] = rules->GetMoveFromList(n.moves, ci);
			edges[ei].prob = prob;
			GameState* ns = rules->ApplyMove(n.state, mv, n.cp);
			const size_t before = nodes.size();
			const int child = make_node(ns);                      // may reallocate `nodes`: n is dead after this line
			edges[ei].child = child;
			if (nodes.size() == before) {                         // transposition into an existing node: keep descending
				if (nodes[child].visit_id == visit_id) {
					charge(p, (uint32_t)cfg.ppl); sims += cfg.ppl;
					backprop(p.path, cfg.cycle_value * cfg.ppl, cfg.cycle_value * cfg.ppl);
					return RESOLVED;
				}
				ni = child; continue;
			}
			if (cfg.expand_all) { ni = child; continue; }         // the new node is fresh: the next iteration makes it the leaf (or resolves a terminal)
			p.npath.push_back(child);
			if (nodes[child].terminal) {
				charge(p, (uint32_t)cfg.ppl); sims += cfg.ppl;
				backprop(p.path, nodes[child].score[0] / 100.0f * cfg.ppl, nodes[child].score[1] / 100.0f * cfg.ppl);
				return RESOLVED;
			}
			p.leaf = child;
			charge(p, (uint32_t)cfg.ppl); sims += cfg.ppl;
			return LEAF;
		}
	}

	// expand_size in the reference-like mode (one playout per simulation): the first `expand_size` positions of the
	// playout below the leaf become permanent nodes as well (expansion, MCTSPlayer.cpp:524-562).  The playout is a pure
	// function of (key, id): the host replays its first moves from the same Philox words the device used.
	void replay_expansion(Pending& p, uint64_t id)
	{
		int ni = p.leaf;
		gai::Philox4 blk{};
		const int steps = std::min(cfg.expand_size, cfg.max_plies);
		for (int t = 0; t < steps; ++t) {
			if (nodes[ni].terminal) break;
			expand(nodes[ni]);
			Node& n = nodes[ni];
			if (n.n_edges == 0) break;
			if ((t & 3) == 0) blk = gai::choice_block(key, id, (uint32_t)t >> 2);
			const int ci = (int)gai::choice_index(blk.v[t & 3], (uint32_t)n.n_edges);
			const int ei = n.first_edge + ci;
			if (edges[ei].child < 0) {
				auto [mv, prob] = rules->GetMoveFromList(n.moves, ci);
				edges[ei].prob = prob;
				const int child = make_node(rules->ApplyMove(n.state, mv, n.cp));        // may reallocate `nodes`
				edges[ei].child = child;
			}
			edges[ei].visits += 1;
			nodes[edges[ei].child].visits += 1;
			p.path.push_back(ei);
			ni = edges[ei].child;
		}
	}

	// A batch: up to cfg.batch selections.  expand_all: the leaves are PARENTS, every successor is played out.
	struct Batch {
		std::vector<Pending> pend; std::vector<gai_loa_state> leaves; std::vector<uint32_t> off; std::vector<gai_loa_result> res;
		int nl = 0, ticket = -1; uint64_t first_id = 0;
		explicit Batch(int b) : pend((size_t)b), leaves((size_t)b), off((size_t)b + 1, 0u) {}
	};
	long fill_batch(int root, Batch& b, long sims_so_far, long sim_limit)
	{
		b.nl = 0; b.off[0] = 0;
		long sims = 0;
		int misses = 0;
		for (int k = 0; k < cfg.batch; ++k) {
			if (sim_limit > 0 && sims_so_far + sims >= sim_limit) break;
			const Sel r = select_leaf(root, b.pend[b.nl], sims);
			if (r == COLLISION) { if (++misses > 16) break; --k; continue; }      // the tree under the root is saturated with pending expansions
			if (r == LEAF) {
				const Node& lf = nodes[b.pend[b.nl].leaf];
				memcpy(&b.leaves[b.nl], rules->GetStateHash(lf.state), 24);
				b.off[b.nl + 1] = b.off[b.nl] + (cfg.expand_all ? (uint32_t)lf.n_edges : 1u);
				++b.nl;
			}
		}
		return sims;
	}
	bool launch(Batch& b)
	{
		b.ticket = -1;
		if (b.nl == 0) return true;
		const uint32_t total = b.off[b.nl];
		if (b.res.size() < total) b.res.resize((size_t)total + total / 2 + 64);
		b.first_id = next_playout_id;
		const int rc = cfg.expand_all
			? gai_loa_rollout_children_submit(ctx, b.leaves.data(), (uint32_t)b.nl, b.off.data(), (uint32_t)cfg.ppl, (uint32_t)cfg.max_plies, key, next_playout_id, b.res.data(), &b.ticket)
			: gai_loa_rollout_submit(ctx, b.leaves.data(), (uint32_t)b.nl, (uint32_t)cfg.ppl, (uint32_t)cfg.max_plies, key, next_playout_id, b.res.data(), &b.ticket);
		if (rc != GAI_OK) { error = gai_last_error(ctx); return false; }
		next_playout_id += (uint64_t)total * cfg.ppl;
		return true;
	}
	bool finish(Batch& b)
	{
		if (b.nl == 0) return true;
		gai_rollout_stats st;
		if (gai_wait(ctx, b.ticket, &st) != GAI_OK) { error = gai_last_error(ctx); return false; }
		playouts += st.playouts; plies += st.plies; gpu_ms += st.total_ms;
		++batches;
		const long mean_plies = st.playouts ? (long)(st.plies / st.playouts) : 0;
		for (int i = 0; i < b.nl; ++i) {
			Pending& p = b.pend[i];
			if (path_size) path_size->insert((long)p.path.size() + mean_plies);
			if (cfg.expand_all) {
				Node& lf = nodes[p.leaf];
				float vw = 0.f, vb = 0.f;
				for (int k = 0; k < lf.n_edges; ++k) {
					const gai_loa_result& r = b.res[b.off[i] + (uint32_t)k];
					const float half = 0.5f * (r.draws + r.capped);
					Edge& e = edges[lf.first_edge + k];
					e.visits += (uint32_t)cfg.ppl; e.value[0] += r.white_wins + half; e.value[1] += r.black_wins + half;
					const float f = (1.0f - cfg.weighted_backprop) + cfg.weighted_backprop * e.prob;
					vw += (r.white_wins + half) * f; vb += (r.black_wins + half) * f;
					capped += r.capped;
				}
				lf.pending = false; lf.expanded = true;
				backprop(p.path, vw, vb);
			} else {
				const gai_loa_result& r = b.res[i];
				const float half = 0.5f * (r.draws + r.capped);
				capped += r.capped;
				if (cfg.ppl == 1 && cfg.expand_size > 0) replay_expansion(p, b.first_id + (uint64_t)i);
				backprop(p.path, r.white_wins + half, r.black_wins + half);
			}
		}
		b.nl = 0;
		return true;
	}
	// Run simulations from `root` until the limit with up to `pipeline` batches queued on the context: while the oldest
	// plays out, the others wait behind it on the stream or are being selected here.
	void search(int root, Clock::time_point deadline, long sim_limit)
	{
		const int K = std::max(1, std::min(cfg.pipeline, (int)GAI_MAX_INFLIGHT));
		std::vector<Batch> ring; ring.reserve((size_t)K);
		for (int i = 0; i < K; ++i) ring.emplace_back(cfg.batch);
		std::vector<Batch*> idle, fly;                  // fly: oldest first
		for (auto& b : ring) idle.push_back(&b);
		long sims = 0;
		auto more = [&] { return sim_limit > 0 ? sims < sim_limit : Clock::now() < deadline; };
		for (;;) {
			while (!idle.empty() && more()) {
				Batch* b = idle.back();
				const long got = fill_batch(root, *b, sims, sim_limit);
				sims += got;
				if (b->nl == 0) {
					if (got == 0) break;                 // only collisions: wait for a batch in flight
					continue;                            // everything resolved on the host (terminal nodes, cycles)
				}
				idle.pop_back();
				if (!launch(*b)) return;
				fly.push_back(b);
			}
			if (fly.empty()) break;
			Batch* b = fly.front(); fly.erase(fly.begin());
			if (!finish(*b)) return;
			idle.push_back(b);
		}
	}
};

const char* const KNOWN_KEYS[] = {
	// the reference's (MCplayer_factory.cpp:37-76, GameLauncher.cpp:121, GameController.cpp:152-155)
	"name", "fullname", "provider", "type", "number_of_players", "random_seed", "playout_depth", "expand_size", "explore_exploit_ratio",
	"trace_move_filename", "game_tree_filename", "out_dir", "expand_from_last_permanent_node", "weighted_backprop", "eval_function",
	"best_move_value_eps", "cycle_score", "move_time_limit", "move_sim_limit", "trace", "knows_complete_game_state",
	// this plugin's
	"gpu_batch", "playouts_per_leaf", "expand_all", "pipeline_depth", "devices", "trees_per_gpu", "philox_seed",
};

struct GpuMctsPlayer : IGamePlayer
{
	Settings cfg;
	IGameRules* rules = nullptr;
	std::vector<Tree> trees;                 // one per device
	std::vector<int> devices;
	std::mt19937 tie_rng;                    // final-move tie break: identical on every rank
	Histogram<long> runs_per_move, branching, path_size, node_pool_usage;
	Average playouts_per_sec, plies_per_playout, gpu_busy, find_root, terminal_found, best_is_most_visited, allreduce_us, selections_per_sec;
	double total_batches = 0, collisions = 0;
	int world = 1, rank = 0; bool comm_ready = false;
	int move_nbr = 0, game_nbr = 0;
	std::string last_root;                   // "visits:value_white:value_black,..." of the last move decision (merged over trees/ranks)
	double gc_runs = 0, gc_freed = 0, reused_visits = 0;
	std::ofstream trace_file;
	std::string id_file;

	GpuMctsPlayer(const Settings& s, const std::vector<int>& devs) : cfg(s), devices(devs), tie_rng(s.seed)
	{
		runs_per_move.Bucket(1000); path_size.Bucket(10); node_pool_usage.Bucket(1000);
		if (!cfg.trace.empty()) {
			const std::string path = cfg.out_dir.empty() || cfg.trace.find('/') != std::string::npos ? cfg.trace : cfg.out_dir + "/" + cfg.trace;
			trace_file.open(path, std::ios::app);
			if (!trace_file) std::fprintf(stderr, "gpumctsplayer: cannot open trace file %s (tracing disabled)\n", path.c_str());
		}
	}
	~GpuMctsPlayer() override
	{
		for (auto& t : trees) { if (rules) t.clear(); if (t.ctx) gai_destroy(t.ctx); }
		if (rank == 0 && !id_file.empty()) unlink(id_file.c_str());
	}

	void release() override { delete this; }
	std::string getName() override { return cfg.name; }
	void resetStats() override {}
	void startNewGame(GameState*) override { move_nbr = 0; ++game_nbr; }
	void endGame(int, GameResult) override
	{
		size_t nn = 0;
		for (auto& t : trees) { nn += t.nodes.size(); t.clear(); }
		node_pool_usage.insert((long)nn);
	}

	void setGameRules(IGameRules* gr) override
	{
		rules = gr;
		if (gr->GetStateHashSize() != 6) throw "gpumctsplayer: the GPU rollout path is built for the Lines of Action rules (24-byte state)";
		if (const char* ws = getenv("WORLD_SIZE")) world = atoi(ws);
		if (const char* rk = getenv("RANK")) rank = atoi(rk);
		trees.resize(devices.size());
		for (size_t i = 0; i < devices.size(); ++i) {
			Tree& t = trees[i];
			t.rules = gr; t.cfg = cfg; t.rng.seed(cfg.seed + 7919u * (unsigned)(i + (size_t)rank * devices.size()));
			t.key = cfg.philox_key ^ (0x9E3779B97F4A7C15ull * (uint64_t)(i + (size_t)rank * devices.size()));   // own Philox stream per tree
			if (i == 0) t.path_size = &path_size;
			if (gai_create(devices[i], &t.ctx) != GAI_OK) {
				static std::string msg; msg = std::string("gpumctsplayer: ") + gai_last_error(nullptr);
				throw msg.c_str();                             // no CPU fallback: the player cannot play without its GPU
			}
		}
		// GAI_FORCE_COMM: a single rank still goes through the NCCL communicator (lets a one-GPU box exercise the merge path)
		if (world > 1 || getenv("GAI_FORCE_COMM")) init_comm();
	}
	// rank 0 writes the NCCL unique id to GAI_NCCL_ID_FILE, the other ranks poll for it.  A file left over from a
	// crashed run is not accepted: rank 0 removes it before writing (and when it exits), the others only take a file
	// written in the last two minutes.
	void init_comm()
	{
		const char* path = getenv("GAI_NCCL_ID_FILE");
		if (!path) throw "gpumctsplayer: WORLD_SIZE > 1 (or GAI_FORCE_COMM) needs GAI_NCCL_ID_FILE";
		id_file = path;
		uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES];
		if (rank == 0) {
			unlink(path);
			if (gai_comm_unique_id(id) != GAI_OK) throw "gpumctsplayer: ncclGetUniqueId failed";
			std::string tmp = std::string(path) + ".tmp";
			{ std::ofstream f(tmp, std::ios::binary); f.write((const char*)id, sizeof id); }
			rename(tmp.c_str(), path);
		} else {
			for (int tries = 0;; ++tries) {
				struct stat sb;
				if (stat(path, &sb) == 0 && time(nullptr) - sb.st_mtime < 120) {
					std::ifstream f(path, std::ios::binary);
					if (f && f.read((char*)id, sizeof id)) break;
				}
				if (tries > 1200) throw "gpumctsplayer: timed out waiting for the NCCL id file";
				std::this_thread::sleep_for(std::chrono::milliseconds(100));
			}
		}
		if (gai_comm_init(trees[0].ctx, rank, world, id) != GAI_OK) throw "gpumctsplayer: gai_comm_init failed";
		comm_ready = true;
		// the final-move tie break must be the same stream on every rank: take rank 0's seed
		uint64_t s = rank == 0 ? (uint64_t)cfg.seed : 0ull;
		if (gai_allreduce_u64(trees[0].ctx, &s, 1) != GAI_OK) throw "gpumctsplayer: seed broadcast failed";
		tie_rng.seed((unsigned)s);
	}

	MoveList* selectMove(GameState* pks) override
	{
		if (rules->GetCurrentPlayer(pks) != cfg.player) return rules->GetPlayerLegalMoves(pks, cfg.player);     // the shared noop list
		++move_nbr;
		const auto t0 = Clock::now();
		const auto deadline = t0 + std::chrono::microseconds((long)(cfg.time_limit * 1e6));
		std::vector<int> roots(trees.size());
		for (size_t i = 0; i < trees.size(); ++i) {
			Tree& t = trees[i];
			const size_t before_n = t.nodes.size();
			roots[i] = t.make_node(rules->CopyGameState(pks));                        // reuses the subtree when the state is known
			if (i == 0 && move_nbr > 1) find_root.insert(t.nodes.size() == before_n ? 1.0 : 0.0);      // find_root_node_ratio (MCTSPlayer.cpp:160)
			if (t.nodes.size() > 200000) { const size_t before = t.nodes.size(); roots[i] = t.collect(roots[i]); gc_freed += before - t.nodes.size(); ++gc_runs; }
			t.playouts = t.plies = t.batches = t.capped = t.selections = t.collisions = 0; t.gpu_ms = 0;
			reused_visits += t.nodes[roots[i]].visits;                               // > 0 when the position was found in the old tree
		}
		// a root without moves (terminal position, or no legal move): hand back the rules' own list for this player
		{
			Tree& t0r = trees[0];
			Node& r = t0r.nodes[roots[0]];
			if (!r.terminal) t0r.expand(t0r.nodes[roots[0]]);
			if (t0r.nodes[roots[0]].terminal || t0r.nodes[roots[0]].n_edges == 0) return rules->GetPlayerLegalMoves(pks, cfg.player);
		}
		if (trees.size() == 1) trees[0].search(roots[0], deadline, cfg.sim_limit);
		else {
			std::vector<std::thread> th;
			for (size_t i = 0; i < trees.size(); ++i) th.emplace_back([&, i] { trees[i].search(roots[i], deadline, cfg.sim_limit > 0 ? cfg.sim_limit / (long)trees.size() : -1); });
			for (auto& x : th) x.join();
		}
		for (auto& t : trees) if (!t.error.empty()) { static std::string msg; msg = "gpumctsplayer: " + t.error; throw msg.c_str(); }

		// merge the root statistics of all trees (children are index aligned: the move list order is deterministic)
		for (size_t i = 0; i < trees.size(); ++i) trees[i].expand(trees[i].nodes[roots[i]]);
		const Node& r0 = trees[0].nodes[roots[0]];
		const int nm = r0.n_edges;
		// fixed 96 + 4 slots whatever nm is: ranks that disagree about the position must not disagree about the NCCL count
		const int SLOTS = GAI_LOA_MAX_MOVES + 4;
		std::vector<uint32_t> visits((size_t)SLOTS, 0); std::vector<float> values((size_t)SLOTS * 2, 0.f);
		uint64_t playouts = 0, plies = 0, batches = 0, capped = 0, sel = 0; double gpu_ms = 0;
		for (size_t i = 0; i < trees.size(); ++i) {
			const Node& r = trees[i].nodes[roots[i]];
			for (int m = 0; m < nm && m < r.n_edges; ++m) {
				const Edge& e = trees[i].edges[r.first_edge + m];
				visits[m] += e.visits; values[2 * m] += e.value[0]; values[2 * m + 1] += e.value[1];
			}
			playouts += trees[i].playouts; plies += trees[i].plies; batches += trees[i].batches; capped += trees[i].capped; sel += trees[i].selections;
			collisions += trees[i].collisions; gpu_ms = std::max(gpu_ms, trees[i].gpu_ms);
		}
		if (comm_ready) {
			// every rank must be looking at the same position with the same move list: sum a checksum and its complement
			const Key k = trees[0].key_of(r0.state);
			const uint64_t h = (k.w * 0x9E3779B97F4A7C15ull) ^ (k.b * 0xC2B2AE3D27D4EB4Full) ^ (k.cp << 7) ^ ((uint64_t)nm << 32);
			const uint32_t hl = (uint32_t)h ^ (uint32_t)(h >> 32);
			visits[96] = hl & 0xffffu; visits[97] = 0xffffu - (hl & 0xffffu); visits[98] = hl >> 16; visits[99] = 0xffffu - (hl >> 16);
			const auto ta = Clock::now();
			if (gai_allreduce_root(trees[0].ctx, visits.data(), values.data(), (uint32_t)SLOTS) != GAI_OK) throw "gpumctsplayer: gai_allreduce_root failed";
			allreduce_us.insert(std::chrono::duration<double, std::micro>(Clock::now() - ta).count());
			const uint32_t w = (uint32_t)world;
			if (visits[96] != w * (hl & 0xffffu) || visits[97] != w * (0xffffu - (hl & 0xffffu)) || visits[98] != w * (hl >> 16) || visits[99] != w * (0xffffu - (hl >> 16)))
				throw "gpumctsplayer: the ranks of a root-parallel search are not looking at the same position (seeds / opponents differ between ranks?)";
		}

		// final move: best value, uniformly among those within eps (MCTSPlayer.cpp:601-618)
		std::vector<int> best; float bestv = 0;
		for (int m = 0; m < nm; ++m) {
			const float v = values[2 * m + cfg.player] * trees[0].edges[r0.first_edge + m].prob / (1.0f + visits[m]);
			if (best.empty()) { best.push_back(m); bestv = v; continue; }
			const float diff = v - bestv;
			if (std::fabs(diff) < cfg.eps) best.push_back(m);
			else if (diff > 0) { best.clear(); best.push_back(m); bestv = v; }
		}
		const int pick = best.empty() ? 0 : best[best.size() == 1 ? 0 : std::uniform_int_distribution<size_t>(0, best.size() - 1)(tie_rng)];
		{
			bool most = true;
			for (int m = 0; m < nm; ++m) if (m != pick && visits[pick] < visits[m]) { most = false; break; }
			best_is_most_visited.insert(most ? 1.0 : 0.0);                        // MCTSPlayer.cpp:620-630
			std::ostringstream rs;
			for (int m = 0; m < nm; ++m) rs << (m ? "," : "") << visits[m] << ":" << values[2 * m] << ":" << values[2 * m + 1];
			last_root = rs.str();
		}
		const double secs = std::chrono::duration<double>(Clock::now() - t0).count();
		runs_per_move.insert((long)playouts);
		branching.insert(nm);
		if (secs > 0) playouts_per_sec.insert(playouts / secs);
		if (secs > 0) selections_per_sec.insert(sel / secs);
		if (playouts) { plies_per_playout.insert((double)plies / playouts); terminal_found.insert(1.0 - (double)capped / playouts); }
		if (secs > 0) gpu_busy.insert(gpu_ms * 1e-3 / secs);
		total_batches += batches;
		if (trace_file) {
			auto [mv, pr] = rules->GetMoveFromList(r0.moves, pick); (void)pr;
			trace_file << "game " << game_nbr << " move " << move_nbr << " playouts " << playouts << " seconds " << secs << " chosen " << rules->ToString(mv)
			           << " value " << bestv << " root " << last_root << "\n";
			trace_file.flush();
		}
		return rules->SelectMoveFromList(r0.moves, pick);
	}

	NamedMetrics_t getGameStats() override
	{
		NamedMetrics_t nm;
		// the reference's names (MCTSPlayer.cpp:149-165)
		nm["node_pool_usage"] = node_pool_usage;
		nm["runs_per_move"] = runs_per_move;
		nm["branching_factor"] = branching;
		nm["path_size"] = path_size;
		nm["find_root_node_ratio"] = find_root;
		nm["terminal_node_found_ratio"] = terminal_found;
		nm["best_move_is_most_visited_ratio"] = best_is_most_visited;
		// this plugin's
		nm["playouts_per_sec"] = playouts_per_sec;
		nm["selections_per_sec"] = selections_per_sec;
		nm["plies_per_playout"] = plies_per_playout;
		nm["gpu_busy_fraction"] = gpu_busy;
		nm["gpu_batches"] = total_batches;
		nm["pending_collisions"] = collisions;
		nm["allreduce_us_per_move"] = allreduce_us;
		nm["last_root_stats"] = last_root;
		nm["tree_gc_runs"] = gc_runs; nm["tree_gc_freed_nodes"] = gc_freed; nm["reused_root_visits"] = reused_visits;
		nm["expand_size_effective"] = (double)((!cfg.expand_all && cfg.ppl == 1) ? cfg.expand_size : 0);
		return nm;
	}
};

} // namespace

extern "C" IGamePlayer* createPlayer(int player_number, const PlayerConfig_t* pc)
{
	static std::string err;
	for (auto& kv : pc->attrs) {
		if (!kv.first.empty() && kv.first[0] == '_') continue;             // commented-out attribute (run_config_loa.xml: _move_sim_limit="50")
		bool ok = false;
		for (const char* k : KNOWN_KEYS) if (kv.first == k) { ok = true; break; }
		if (!ok) { err = "gpumctsplayer: unknown config key '" + kv.first + "'"; throw err.c_str(); }
	}
	const std::string type = pc->get_optional<std::string>("type").get_value_or("mcts");
	if (type != "mcts") { err = "gpumctsplayer: player type '" + type + "' is not provided by this plugin (only type=\"mcts\"; MCRLplayer.cpp has no rollout phase to accelerate)"; throw err.c_str(); }
	Settings s;
	s.player = player_number;
	s.seed = pc->get_optional<unsigned>("random_seed").get_value_or((unsigned)Clock::now().time_since_epoch().count());
	s.num_players = pc->get_optional<int>("number_of_players").get_value_or(2);
	s.max_plies = pc->get_optional<int>("playout_depth").get_value_or(20);
	s.C = pc->get_optional<float>("explore_exploit_ratio").get_value_or(1.0f);
	s.eps = pc->get_optional<float>("best_move_value_eps").get_value_or(0.005f);
	s.expand_size = std::max(0, pc->get_optional<int>("expand_size").get_value_or(1));
	s.cycle_value = pc->get_optional<int>("cycle_score").get_value_or(50) / 100.0f;
	s.weighted_backprop = pc->get_optional<float>("weighted_backprop").get_value_or(0.f);
	s.eval_function = pc->get_optional<std::string>("eval_function").get_value_or("");
	s.trace = pc->get_optional<std::string>("trace").get_value_or("");
	s.out_dir = pc->get_optional<std::string>("out_dir").get_value_or("");
	s.batch = std::max(1, pc->get_optional<int>("gpu_batch").get_value_or(1024));
	s.ppl = std::max(1, pc->get_optional<int>("playouts_per_leaf").get_value_or(32));
	s.expand_all = pc->get_optional<int>("expand_all").get_value_or(1) != 0;
	s.pipeline = std::max(1, std::min((int)GAI_MAX_INFLIGHT, pc->get_optional<int>("pipeline_depth").get_value_or(2)));
	s.philox_key = pc->get_optional<uint64_t>("philox_seed").get_value_or((uint64_t)s.seed * 0x9E3779B97F4A7C15ull + 1);
	auto sl = pc->get_optional<long>("move_sim_limit");
	if (sl) s.sim_limit = sl.get(); else s.time_limit = pc->get_optional<double>("move_time_limit").get_value_or(1.0);
	s.name = pc->get_optional<std::string>("fullname").get_value_or(pc->get_optional<std::string>("name").get_value_or("gpumcts"));
	// Keys that are honoured only in part say so once, loudly, instead of being dropped:
	//  * eval_function: the depth-capped playout backs up the rules' evaluation.  For LOA that is the constant 50/50
	//    whatever the name (LinesOfActionZasady.cpp:382-397), which is what a capped GPU playout returns.
	//  * expand_size: exact in the reference-like mode (expand_all=0 playouts_per_leaf=1); in the batched modes the playout
	//    path stays on the device and each selection adds one node (expand_all: one node with statistics for every move).
	if (s.expand_size > 0 && (s.expand_all || s.ppl != 1) && pc->get_optional<int>("expand_size"))
		std::fprintf(stderr, "gpumctsplayer: expand_size=%d adds playout-path nodes only with expand_all=0 playouts_per_leaf=1; this configuration (expand_all=%d playouts_per_leaf=%d) adds one node per selection\n",
		             s.expand_size, s.expand_all ? 1 : 0, s.ppl);
	std::vector<int> devs;
	{
		std::string d = pc->get_optional<std::string>("devices").get_value_or("");
		if (d.empty()) { const char* lr = getenv("LOCAL_RANK"); devs.push_back(lr ? atoi(lr) : 0); }
		else { std::replace(d.begin(), d.end(), ',', ' '); std::istringstream ss(d); int x; while (ss >> x) devs.push_back(x); }
		if (devs.empty()) devs.push_back(0);
		// trees_per_gpu=K: K independent trees (own host thread, context and stream each) on every listed GPU, merged at the
		// root like trees on different GPUs.  One tree keeps a GPU ~60-75 % busy (its single selection thread is the limit);
		// two interleave their batches on the device: 1.65e8 playouts/s at 90 % busy against 1.0-1.2e8
		// (profiles/r02_player_trees_per_gpu.jsonl)
		const int k = std::max(1, std::min(8, pc->get_optional<int>("trees_per_gpu").get_value_or(1)));
		const std::vector<int> base = devs;
		for (int r = 1; r < k; ++r) devs.insert(devs.end(), base.begin(), base.end());
	}
	return new GpuMctsPlayer(s, devs);
}
