// tests/helpers/host_rules.cu -- TEST HELPER: runs the __host__ __device__ rule functions of the product headers on the
// CPU so that `pytest -m "not gpu"` can compare them with the oracle before any GPU time is spent.
#include "pan_rules.cuh"
#include "loa_fast.cuh"
#include "loa_sweep.cuh"
#include "gameai/philox.h"
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
unsigned host_loa_slt_entry(unsigned i) { return loa::slt_entry(i); }
unsigned host_loa_tpc_entry(unsigned b) { return loa::tpc_entry(b); }
int host_loa_pos_c(int s) { return loa::pos_c(s); }
int host_loa_pos_d1(int s) { return loa::pos_d1(s); }
int host_loa_pos_d2(int s) { return loa::pos_d2(s); }
unsigned host_loa_phantom_d1(int s) { return loa::phantom_d1(s); }
unsigned host_loa_phantom_d2(int s) { return loa::phantom_d2(s); }

// ---- the static-sweep formulation (loa_sweep.cuh) run on the CPU: count + rank-select of every move, and whole
// playouts through sweep_rows / sweep_select / apply_move_fast (the code the rollout kernel executes per ply)
static const uint16_t* host_lut()
{
	static uint16_t* lut = nullptr;
	if (!lut) {
		lut = new uint16_t[loa::LUT_BYTES / 2]();
		for (unsigned e = 0; e < 256; ++e) for (unsigned o = 0; o < 256; ++o) lut[o + e * loa::LUT_STRIDE] = loa::lut_entry(o, e);
	}
	return lut;
}
static loa::SweepTabs host_tabs()
{
	static loa::u32 slt[256], tpc[256]; static bool init = false;
	if (!init) { for (unsigned i = 0; i < 256; ++i) { slt[i] = loa::slt_entry(i); tpc[i] = loa::tpc_entry(i); } init = true; }
	return loa::sweep_tabs_host(host_lut(), slt, tpc);
}
static const loa::SquareTable g_sqt = loa::make_square_table();
static const loa::DiagTables g_diag = loa::make_diag_tables();
static const loa::LineTable g_lines = loa::make_line_table();
static const loa::NbxTable g_nbx = loa::make_nbx_table();
static const loa::ChiTable g_chi = loa::make_chi_table();
// Euler characteristic (components - holes) of a board by the product's table-driven code
int host_loa_chi(uint64_t b) { return loa::chi_of(b, g_nbx.w, g_chi.d); }
void host_loa_sweep_movegen_batch(const uint64_t* states, unsigned n, uint8_t* counts, uint8_t* moves)
{
	const loa::SweepTabs lut = host_tabs();
	for (unsigned i = 0; i < n; ++i) {
		const uint64_t w = states[3 * i], b = states[3 * i + 1]; const unsigned side = (unsigned)states[3 * i + 2] & 1u;
		const loa::Side my = loa::make_side(side ? b : w, g_sqt.w), en = loa::make_side(side ? w : b, g_sqt.w);
		const loa::RowCounts rc = loa::sweep_rows(my, en, lut);
		counts[i] = (uint8_t)rc.n;
		for (unsigned k = 0; k < rc.n && k < 96; ++k) {
			int from, dir;
			loa::sweep_select(rc, k, lut, &from, &dir);
			moves[(size_t)i * 192 + 2 * k] = (uint8_t)from;
			moves[(size_t)i * 192 + 2 * k + 1] = (uint8_t)loa::move_target(my.R | en.R, from, dir, g_diag.ld, g_diag.rd);
		}
	}
}
void host_loa_sweep_playout_batch(const uint64_t* states, unsigned n, unsigned max_plies, uint64_t key, uint64_t first_id,
                                  uint8_t* outcome, uint32_t* plies, uint64_t* final_states)
{
	const loa::SweepTabs lut = host_tabs();
	for (unsigned i = 0; i < n; ++i) {
		unsigned side = (unsigned)states[3 * i + 2] & 1u;
		loa::Side my = loa::make_side(side ? states[3 * i + 1] : states[3 * i], g_sqt.w), en = loa::make_side(side ? states[3 * i] : states[3 * i + 1], g_sqt.w);
		unsigned ply = 0; int out = -1;
		gai::Philox4 rnd{};
		int chi_my = loa::chi_of(my.R, g_nbx.w, g_chi.d), chi_en = loa::chi_of(en.R, g_nbx.w, g_chi.d);
		for (;;) {
			// the kernel's test (wstep_sweep): the flood fill only for a colour whose chi is at most 1
			const bool c_en = chi_en <= 1 && loa::connected(en.R), c_my = chi_my <= 1 && loa::connected(my.R);
			if (c_en | c_my) { out = loa::outcome_of(side ? c_en : c_my, side ? c_my : c_en); break; }
			if (ply >= max_plies) { out = loa::OUT_CAPPED; break; }
			const loa::RowCounts rc = loa::sweep_rows(my, en, lut);
			if ((ply & 3u) == 0) rnd = gai::choice_block(key, first_id + i, ply >> 2);
			if (rc.n) {
				const unsigned idx = gai::choice_index(rnd.v[ply & 3u], rc.n);
				int from, dir;
				loa::sweep_select(rc, idx, lut, &from, &dir);
				const int to = loa::move_target_tab(my.R | en.R, from, dir, &g_lines.line[0][0]);
				if (to != loa::move_target(my.R | en.R, from, dir, g_diag.ld, g_diag.rd)) { out = 0xee; break; }
				loa::apply_move_chi(my, en, chi_my, chi_en, from, to, g_sqt.w, g_nbx.w, g_chi.d);
				// the incrementally maintained chi must equal a recount from scratch
				if (chi_my != loa::chi_of(my.R, g_nbx.w, g_chi.d) || chi_en != loa::chi_of(en.R, g_nbx.w, g_chi.d)) { out = 0xed; break; }
			}
			const loa::Side t = my; my = en; en = t;
			const int tc = chi_my; chi_my = chi_en; chi_en = tc;
			side ^= 1u; ++ply;
		}
		outcome[i] = (uint8_t)out; plies[i] = ply;
		final_states[3 * i] = side ? en.R : my.R; final_states[3 * i + 1] = side ? my.R : en.R; final_states[3 * i + 2] = side;
		// the rotations must still describe the same position
		const loa::Side cm = loa::make_side(my.R, g_sqt.w), ce = loa::make_side(en.R, g_sqt.w);
		if (cm.C != my.C || cm.D1 != my.D1 || cm.D2 != my.D2 || ce.C != en.C || ce.D1 != en.D1 || ce.D2 != en.D2) outcome[i] = 0xee;
	}
}
}
