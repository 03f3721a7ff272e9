// tests/helpers/fake_gpu.cpp -- TEST HELPER, never shipped, never linked: LD_PRELOAD-ed by tests/test_host.py in front of
// libgameai_gpu.so so that the HOST logic of the GPU MCTS player (selection with virtual loss, lazy expansion, the two
// pipelined leaf batches, transposition table, tree reuse + garbage collection, final-move rule) can be exercised by
// `pytest -m "not gpu"`.  It plays no games: gai_loa_rollout_async "returns" a deterministic pseudo-random split of the
// playouts per leaf (a hash of the leaf's position, biased towards the side with more mobility-irrelevant material --
// here simply towards White having fewer pieces to connect).  The product has no CPU path; this file is not one.
#include "gameai_gpu.h"
#include <cstdint>
#include <cstring>

struct gai_ctx { const gai_loa_state* leaves; uint32_t nl, ppl; gai_loa_result* out; uint64_t launches; gai_loa_state copy[1 << 16]; };
static uint64_t mix(uint64_t h) { h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31; return h; }
extern "C" {
int gai_create(int, gai_ctx** o) { *o = new gai_ctx(); (*o)->launches = 0; return GAI_OK; }
void gai_destroy(gai_ctx* c) { delete c; }
const char* gai_last_error(const gai_ctx*) { return "fake_gpu (test helper)"; }
int gai_loa_rollout_async(gai_ctx* c, const gai_loa_state* leaves, uint32_t n, uint32_t ppl, uint32_t, uint64_t, uint64_t, gai_loa_result* out)
{
	if (n > (1u << 16)) return GAI_ERR_INVALID;
	memcpy(c->copy, leaves, (size_t)n * sizeof(gai_loa_state));      // like the real call: the caller may reuse `leaves` at once
	c->nl = n; c->ppl = ppl; c->out = out; c->launches += 1;
	return GAI_OK;
}
int gai_sync(gai_ctx* c)
{
	for (uint32_t i = 0; i < c->nl; ++i) {
		const uint64_t* w = reinterpret_cast<const uint64_t*>(&c->copy[i]);
		const int nw = __builtin_popcountll(w[0]), nb = __builtin_popcountll(w[1]);
		const uint64_t h = mix(w[0] ^ mix(w[1] ^ (w[2] & 1)));
		uint32_t ww = (uint32_t)((h & 0xffff) * (c->ppl + 1) >> 16);                  // uniform in [0, ppl]
		if (nw < nb && ww < c->ppl) ww += (uint32_t)((h >> 20) % (c->ppl - ww + 1)) / 2;   // fewer white pieces: a little better for White
		c->out[i].white_wins = ww; c->out[i].black_wins = c->ppl - ww; c->out[i].draws = 0; c->out[i].capped = 0;
	}
	c->nl = 0;
	return GAI_OK;
}
int gai_last_rollout_stats(gai_ctx* c, gai_rollout_stats* s) { memset(s, 0, sizeof *s); s->playouts = (uint64_t)c->ppl * 1024; s->plies = s->playouts * 100; s->kernel_ms = s->total_ms = 0.1f; return GAI_OK; }
}
