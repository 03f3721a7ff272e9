// tests/helpers/host_rules.cu -- TEST HELPER: runs the __host__ __device__ rule functions of the product headers on the
// CPU so that `pytest -m "not gpu"` can compare them with the oracle before any GPU time is spent.
#include "pan_rules.cuh"
#include "loa_fast.cuh"
#include <string.h>

extern "C" {
int host_pan_movegen(int np, const uint64_t* st, uint64_t* moves)
{
	pan::State s;
	if (!pan::load_state(s, (const pan::u64*)st)) return -1;
	(void)np;
	pan::Moves m = pan::gen_moves(s);
	if ((st[0] >> 51) & 1ull) m.n = 0;
	for (pan::u32 k = 0; k < m.n; ++k) { pan::u32 c, cnt, op; pan::nth_move(m, k, &c, &cnt, &op); moves[k] = pan::move_word(c, cnt, op); }
	return (int)m.n;
}
int host_pan_apply(int np, const uint64_t* st, uint64_t mv, uint64_t* out)
{
	pan::State s;
	if (!pan::load_state(s, (const pan::u64*)st) || !pan::is_crisp(mv & pan::CARDS48)) return -1;
	pan::apply(s, np, pan::squeeze(mv & pan::CARDS48), (pan::u32)((mv >> 48) & 7u), (pan::u32)((mv >> 51) & 7u));
	pan::store_state((pan::u64*)out, s, np);
	return 0;
}
int host_pan_terminal(int np, const uint64_t* st, int* score, int* ev0, int* ev1)
{
	pan::State s;
	if (!pan::load_state(s, (const pan::u64*)st)) return -1;
	pan::score(s, np, score);
	for (int p = 0; p < np; ++p) { ev0[p] = pan::eval_one(s, 0, p); ev1[p] = pan::eval_one(s, 1, p); }
	return pan::is_terminal(s, np) ? 1 : 0;
}
// LOA line LUT: entry for own/enemy occupancy bytes, and the per-square rotation constants
int host_loa_lut_entry(unsigned o, unsigned e) { return loa::lut_entry(o, e); }
int host_loa_pos_c(int s) { return loa::pos_c(s); }
int host_loa_pos_d1(int s) { return loa::pos_d1(s); }
int host_loa_pos_d2(int s) { return loa::pos_d2(s); }
unsigned host_loa_phantom_d1(int s) { return loa::phantom_d1(s); }
unsigned host_loa_phantom_d2(int s) { return loa::phantom_d2(s); }
}
