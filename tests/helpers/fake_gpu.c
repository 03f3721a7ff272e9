/* tests/helpers/fake_gpu.c -- TEST INFRASTRUCTURE ONLY: a stand-in for libgameai_gpu.so that plays the batches on the
 * CPU through the oracle (oracle/libloa_oracle.so).  It exists so that the HOST logic of the GPU-backed player
 * (selection, virtual loss, batch-by-expansion, pipelining, back-propagation, config keys, statistics) can be exercised by
 * `pytest -m "not gpu"` without a device: tests LD_PRELOAD it under GameLauncher.  Nothing in the product loads it; the
 * real library has no CPU path (tests/test_abi.py::test_no_cpu_fallback).
 * GAI_FAKE_MODE=hash: no games are played at all -- every leaf "returns" a deterministic pseudo-random split of its
 * playouts (a hash of the position), for tests that need many simulations (tree growth, garbage collection) quickly. */
#include "gameai_gpu.h"
#include "../../oracle/loa_oracle.h"
#include <stdlib.h>
#include <string.h>

struct gai_ctx { char err[256]; uint64_t playouts[GAI_MAX_INFLIGHT], plies[GAI_MAX_INFLIGHT]; int busy[GAI_MAX_INFLIGHT]; unsigned gen[GAI_MAX_INFLIGHT]; int bad[GAI_MAX_INFLIGHT]; uint64_t launches; };

int gai_abi_version(void) { return GAI_ABI_VERSION; }
int gai_create(int device, gai_ctx** out) { (void)device; *out = (gai_ctx*)calloc(1, sizeof(gai_ctx)); return GAI_OK; }
void gai_destroy(gai_ctx* c) { free(c); }
const char* gai_last_error(const gai_ctx* c) { return c ? c->err : "fake"; }
int gai_sync(gai_ctx* c) { (void)c; return GAI_OK; }
uint64_t gai_launch_count(const gai_ctx* c) { return c->launches; }

static int slot_of(gai_ctx* c) { for (int i = 0; i < GAI_MAX_INFLIGHT; ++i) if (!c->busy[i]) return i; return -1; }

static uint64_t mix(uint64_t h) { h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31; return h; }
static void play(gai_ctx* c, int s, const uint64_t* st, uint32_t n, uint32_t ppl, uint32_t max_plies, uint64_t key, uint64_t first_id, gai_loa_result* out)
{
	const char* mode = getenv("GAI_FAKE_MODE");
	if (mode && !strcmp(mode, "hash")) {
		for (uint32_t i = 0; i < n; ++i) {
			const uint64_t h = mix(st[3 * i] ^ mix(st[3 * i + 1] ^ (st[3 * i + 2] & 1)));
			const uint32_t ww = (uint32_t)(((h & 0xffff) * (ppl + 1)) >> 16);            /* uniform in [0, ppl] */
			out[i].white_wins = ww; out[i].black_wins = ppl - ww; out[i].draws = 0; out[i].capped = 0;
		}
		c->playouts[s] = (uint64_t)n * ppl; c->plies[s] = c->playouts[s] * 100; c->launches += 3;
		return;
	}
	uint64_t* w = (uint64_t*)malloc(8 * (size_t)(n ? n : 1)); uint64_t* b = (uint64_t*)malloc(8 * (size_t)(n ? n : 1)); uint8_t* cp = (uint8_t*)malloc(n ? n : 1);
	for (uint32_t i = 0; i < n; ++i) { w[i] = st[3 * i]; b[i] = st[3 * i + 1]; cp[i] = (uint8_t)(st[3 * i + 2] & 1); }
	memset(out, 0, sizeof(gai_loa_result) * (size_t)n);
	c->plies[s] = loa_oracle_playout_batch(w, b, cp, n, ppl, max_plies, key, first_id, (uint32_t*)out);
	c->playouts[s] = (uint64_t)n * ppl;
	free(w); free(b); free(cp);
	c->launches += 3;
}

int gai_loa_rollout_submit(gai_ctx* c, const gai_loa_state* leaves, uint32_t n, uint32_t ppl, uint32_t max_plies, uint64_t key, uint64_t first_id, gai_loa_result* out, int* ticket)
{
	const int s = slot_of(c);
	if (s < 0) { strcpy(c->err, "all slots busy"); return GAI_ERR_INVALID; }
	c->bad[s] = 0;
	play(c, s, (const uint64_t*)leaves, n, ppl, max_plies, key, first_id, out);
	c->busy[s] = 1; c->gen[s] += 1; *ticket = (int)(c->gen[s] * GAI_MAX_INFLIGHT + (unsigned)s);
	return GAI_OK;
}

int gai_loa_rollout_children_submit(gai_ctx* c, const gai_loa_state* parents, uint32_t n, const uint32_t* off, uint32_t ppl, uint32_t max_plies, uint64_t key, uint64_t first_id, gai_loa_result* out, int* ticket)
{
	const int s = slot_of(c);
	if (s < 0) { strcpy(c->err, "all slots busy"); return GAI_ERR_INVALID; }
	const uint32_t total = n ? off[n] : 0;
	uint64_t* kids = (uint64_t*)malloc(24 * (size_t)(total ? total : 1));
	c->bad[s] = 0;
	for (uint32_t i = 0; i < n; ++i) {
		uint8_t mv[2 * LOA_MAX_MOVES + 16];
		const uint64_t w = parents[i].white, b = parents[i].black; const int cp = (int)(parents[i].current_player & 1);
		const int cnt = loa_oracle_movegen(w, b, cp, mv);
		if ((uint32_t)cnt != off[i + 1] - off[i]) { c->bad[s] = 1; break; }
		for (int k = 0; k < cnt; ++k) {
			uint64_t ow, ob; int ocp;
			loa_oracle_apply(w, b, cp, mv[2 * k], mv[2 * k + 1], &ow, &ob, &ocp);
			uint64_t* d = kids + 3 * (size_t)(off[i] + (uint32_t)k); d[0] = ow; d[1] = ob; d[2] = (uint64_t)ocp;
		}
	}
	if (!c->bad[s]) play(c, s, kids, total, ppl, max_plies, key, first_id, out);
	free(kids);
	c->busy[s] = 1; c->gen[s] += 1; *ticket = (int)(c->gen[s] * GAI_MAX_INFLIGHT + (unsigned)s);
	return GAI_OK;
}

int gai_wait(gai_ctx* c, int ticket, gai_rollout_stats* st)
{
	const int s = ticket % GAI_MAX_INFLIGHT;
	if (ticket < 0 || !c->busy[s] || c->gen[s] != (unsigned)ticket / GAI_MAX_INFLIGHT) { strcpy(c->err, "no such ticket"); return GAI_ERR_INVALID; }
	c->busy[s] = 0;
	if (st) { st->playouts = c->playouts[s]; st->plies = c->plies[s]; st->kernel_ms = 0.01f; st->total_ms = 0.01f; }
	if (c->bad[s]) { strcpy(c->err, "child count mismatch"); return GAI_ERR_INVALID; }
	return GAI_OK;
}
/* the synchronous host-buffer call (used by the INTEGRATION.md section-2 stub, oracle/ref_mcts_gpu_capi.cpp) */
int gai_loa_rollout(gai_ctx* c, const gai_loa_state* leaves, uint32_t n, uint32_t ppl, uint32_t max_plies, uint64_t key, uint64_t first_id, gai_loa_result* out, gai_rollout_stats* st)
{
	if (st) memset(st, 0, sizeof *st);
	if (n == 0) return GAI_OK;
	if (ppl == 0) { strcpy(c->err, "playouts_per_leaf must be >= 1"); return GAI_ERR_INVALID; }
	play(c, 0, (const uint64_t*)leaves, n, ppl, max_plies, key, first_id, out);
	if (st) { st->playouts = c->playouts[0]; st->plies = c->plies[0]; st->kernel_ms = st->total_ms = 0.01f; }
	return GAI_OK;
}
int gai_inflight(const gai_ctx* c) { int k = 0; for (int i = 0; i < GAI_MAX_INFLIGHT; ++i) k += c->busy[i]; return k; }

/* Pan: determinized rollouts through the Pan oracle (gai_pan_determinize is host-only code of the real library) */
uint64_t pan_oracle_playout_batch(int np, const uint64_t* states, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies, int eval_kind,
                                  uint64_t key, uint64_t first_id, uint32_t* res);
int gai_pan_rollout(gai_ctx* c, const gai_pan_state* leaves, uint32_t n, int np, uint32_t ppl, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                    gai_pan_result* out, gai_rollout_stats* st)
{
	if (st) memset(st, 0, sizeof *st);
	if (n == 0) return GAI_OK;
	memset(out, 0, sizeof(gai_pan_result) * (size_t)n);
	const uint64_t plies = pan_oracle_playout_batch(np, (const uint64_t*)leaves, n, ppl, max_plies, eval_kind, key, first_id, (uint32_t*)out);
	if (st) { st->playouts = (uint64_t)n * ppl; st->plies = plies; st->kernel_ms = st->total_ms = 0.01f; }
	c->launches += 2;
	return GAI_OK;
}

/* single-process stand-ins for the NCCL entry points (world size 1) */
int gai_comm_unique_id(uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES]) { memset(id, 7, GAI_NCCL_UNIQUE_ID_BYTES); return GAI_OK; }
int gai_comm_init(gai_ctx* c, int rank, int world, const uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES]) { (void)c; (void)rank; (void)id; return world == 1 ? GAI_OK : GAI_ERR_NCCL; }
int gai_allreduce_root(gai_ctx* c, uint32_t* v, float* f, uint32_t n) { (void)c; (void)v; (void)f; (void)n; return GAI_OK; }
int gai_allreduce_u64(gai_ctx* c, uint64_t* v, uint32_t n) { (void)c; (void)v; (void)n; return GAI_OK; }
