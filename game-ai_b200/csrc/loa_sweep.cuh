// game-ai_b200/csrc/loa_sweep.cuh -- move COUNT and rank-SELECT by a static sweep over the 32 line bytes (device).
//
// The per-piece formulation (loa_fast.cuh piece_moves_fast) runs, in a warp, for the MAXIMUM number of own pieces over
// its 32 lanes -- ncu: 10.3 iterations at 16 of 32 lanes active, 68 % of all issued instructions.  This header removes
// the data-dependent loop from counting: every lane does exactly the same work per ply.
//
//   1. sweep : for each of the 8 bytes of each rotation (R, C, D1, D2) one or two LUT lookups (two when the byte holds
//              two complementary diagonals) -> a 16-bit word, 2 bits per cell (may the own piece on the cell move
//              towards lower / higher cells).  32 words, byte indices and phantom masks are compile-time constants.
//   2. rows  : per-row move totals.  R words belong to one row each (popcount).  C / D1 / D2 words have one cell per
//              row (D1 after a 16-bit rotation by 2k), so the per-row totals are a POSITIONAL popcount over 24 words:
//              a carry-save adder tree on bit planes (LOP3 only), folded to 6-bit counts and spread to byte lanes with
//              integer multiplies (FMA pipe).  A byte-lane prefix sum gives n and the row that holds move idx.
//   3. piece : inside that row the reference order is by column, then direction: walk the (1-3) own pieces of the row
//              with the per-piece lookup.
// Move ORDER is the reference's (LinesOfActionZasady.cpp:180-223): ascending square, then W E N S SW NE SE NW.
#pragma once
#include "loa_fast.cuh"

namespace loa {

template <int I> __device__ __forceinline__ u32 static_byte(u64 x)
{
	return __byte_perm(I < 4 ? (u32)x : (u32)(x >> 32), 0u, 0x4440u | (u32)(I & 3));
}
__device__ __forceinline__ u32 lut16(const uint16_t* __restrict__ lut, u32 o, u32 e) { return lut[o + e * (u32)LUT_STRIDE]; }

// phantom masks of the two diagonals sharing byte K (compile time)
template <int K> struct D1Ph { static constexpr u32 a = (0xffu << (8 - K)) & 0xffu, b = (1u << (8 - K)) - 1u; static constexpr int len_a = 8 - K, len_b = K; };
template <int K> struct D2Ph { static constexpr u32 a = (0xffu << (K + 1)) & 0xffu, b = (1u << (K + 1)) - 1u; static constexpr int len_a = K + 1, len_b = 7 - K; };

template <int K> __device__ __forceinline__ u32 sweep_r(const Side& my, const Side& en, const uint16_t* lut) { return lut16(lut, static_byte<K>(my.R), static_byte<K>(en.R)); }
template <int K> __device__ __forceinline__ u32 sweep_c(const Side& my, const Side& en, const uint16_t* lut) { return lut16(lut, static_byte<K>(my.C), static_byte<K>(en.C)); }
template <int K, class PH> __device__ __forceinline__ u32 sweep_diag(u64 myb, u64 enb, const uint16_t* lut)
{
	const u32 o = static_byte<K>(myb), e = static_byte<K>(enb);
	u32 v = 0;
	if (PH::len_a >= 2) v = lut16(lut, o | PH::a, e | PH::a);          // a one-cell line has no moves
	if (PH::len_b >= 2) v |= lut16(lut, o | PH::b, e | PH::b);         // the two lines own disjoint cells
	return v;
}
__device__ __forceinline__ u32 rotl16(u32 x, int s)                     // x < 65536, 0 <= s < 16
{
	const u32 d = x * 0x10001u;
	return (d >> ((16 - s) & 15)) & 0xffffu;
}

struct FullAdd { u32 s, c; };
__device__ __forceinline__ FullAdd fa(u32 a, u32 b, u32 c) { FullAdd r; r.s = a ^ b ^ c; r.c = (a & b) | (c & (a ^ b)); return r; }

struct RowCounts {
	u32 plo, phi;      // inclusive prefix sums of the per-row move totals, one byte per row (rows 0-3 / 4-7)
	u32 n;             // total number of legal moves
};

// steps 1 + 2.  Every lane executes exactly this straight-line code.
__device__ __forceinline__ RowCounts sweep_rows(const Side& my, const Side& en, const uint16_t* __restrict__ lut)
{
	// ---- R lines: all cells of a word lie in one row
	const u32 r0 = sweep_r<0>(my, en, lut), r1 = sweep_r<1>(my, en, lut), r2 = sweep_r<2>(my, en, lut), r3 = sweep_r<3>(my, en, lut);
	const u32 r4 = sweep_r<4>(my, en, lut), r5 = sweep_r<5>(my, en, lut), r6 = sweep_r<6>(my, en, lut), r7 = sweep_r<7>(my, en, lut);
	u32 tlo = (u32)__popc(r0) + ((u32)__popc(r1) << 8) + ((u32)__popc(r2) << 16) + ((u32)__popc(r3) << 24);
	u32 thi = (u32)__popc(r4) + ((u32)__popc(r5) << 8) + ((u32)__popc(r6) << 16) + ((u32)__popc(r7) << 24);

	// ---- C, D1, D2 lines: one cell per row; two 16-bit words per register
	u32 w[12];
	w[0] = sweep_c<0>(my, en, lut) | (sweep_c<1>(my, en, lut) << 16);
	w[1] = sweep_c<2>(my, en, lut) | (sweep_c<3>(my, en, lut) << 16);
	w[2] = sweep_c<4>(my, en, lut) | (sweep_c<5>(my, en, lut) << 16);
	w[3] = sweep_c<6>(my, en, lut) | (sweep_c<7>(my, en, lut) << 16);
	// D2: cell = row already
	w[4] = sweep_diag<0, D2Ph<0>>(my.D2, en.D2, lut) | (sweep_diag<1, D2Ph<1>>(my.D2, en.D2, lut) << 16);
	w[5] = sweep_diag<2, D2Ph<2>>(my.D2, en.D2, lut) | (sweep_diag<3, D2Ph<3>>(my.D2, en.D2, lut) << 16);
	w[6] = sweep_diag<4, D2Ph<4>>(my.D2, en.D2, lut) | (sweep_diag<5, D2Ph<5>>(my.D2, en.D2, lut) << 16);
	w[7] = sweep_diag<6, D2Ph<6>>(my.D2, en.D2, lut) | (sweep_diag<7, D2Ph<7>>(my.D2, en.D2, lut) << 16);
	// D1: cell = column c, row = (c + k) & 7  ->  rotate the 2-bit fields left by k cells
	w[8]  = sweep_diag<0, D1Ph<0>>(my.D1, en.D1, lut) | (rotl16(sweep_diag<1, D1Ph<1>>(my.D1, en.D1, lut), 2) << 16);
	w[9]  = rotl16(sweep_diag<2, D1Ph<2>>(my.D1, en.D1, lut), 4) | (rotl16(sweep_diag<3, D1Ph<3>>(my.D1, en.D1, lut), 6) << 16);
	w[10] = rotl16(sweep_diag<4, D1Ph<4>>(my.D1, en.D1, lut), 8) | (rotl16(sweep_diag<5, D1Ph<5>>(my.D1, en.D1, lut), 10) << 16);
	w[11] = rotl16(sweep_diag<6, D1Ph<6>>(my.D1, en.D1, lut), 12) | (rotl16(sweep_diag<7, D1Ph<7>>(my.D1, en.D1, lut), 14) << 16);

	// positional popcount of the 12 registers: carry-save adders on bit planes
	const FullAdd a0 = fa(w[0], w[1], w[2]), a1 = fa(w[3], w[4], w[5]), a2 = fa(w[6], w[7], w[8]), a3 = fa(w[9], w[10], w[11]);
	const FullAdd b0 = fa(a0.s, a1.s, a2.s);
	const u32 p1 = b0.s ^ a3.s, c5 = b0.s & a3.s;                                    // ones
	const FullAdd t0 = fa(a0.c, a1.c, a2.c), t1 = fa(a3.c, b0.c, c5);
	const u32 p2 = t0.s ^ t1.s, d2 = t0.s & t1.s;                                    // twos
	const FullAdd f4 = fa(t0.c, t1.c, d2);
	const u32 p4 = f4.s, p8 = f4.c;                                                  // fours, eights   (count <= 12)
	// fold the two 16-bit halves: (p8 p4 p2 p1)[lo] + (p8 p4 p2 p1)[hi]  -> 5 planes in the low 16 bits
	const u32 h1 = p1 >> 16, h2 = p2 >> 16, h4 = p4 >> 16, h8 = p8 >> 16;
	const u32 s1 = p1 ^ h1, k1 = p1 & h1;
	const FullAdd g2 = fa(p2, h2, k1), g4 = fa(p4, h4, g2.c), g8 = fa(p8, h8, g4.c);
	const u32 q1 = s1, q2 = g2.s, q4 = g4.s, q8 = g8.s, q16 = g8.c;                  // count <= 24 per (row, sense)
	// fold the two senses of a cell (even / odd bit of a field): 6 planes on the even bits
	const u32 M = 0x5555u;
	const u32 e1 = q1 & M, o1 = (q1 >> 1) & M, e2 = q2 & M, o2 = (q2 >> 1) & M, e4 = q4 & M, o4 = (q4 >> 1) & M;
	const u32 e8 = q8 & M, o8 = (q8 >> 1) & M, e16 = q16 & M, o16 = (q16 >> 1) & M;
	const u32 z1 = e1 ^ o1, y1 = e1 & o1;
	const FullAdd z2 = fa(e2, o2, y1), z4 = fa(e4, o4, z2.c), z8 = fa(e8, o8, z4.c), z16 = fa(e16, o16, z8.c);
	// spread plane bits (bit 2r) to byte lanes (byte r) and weight them: all on the FMA pipe
	const u32 SP = 0x00041041u, BM = 0x01010101u;
#define GAI_SPREAD_LO(x) ((((x) & 0x55u) * SP) & BM)
#define GAI_SPREAD_HI(x) (((((x) >> 8) & 0x55u) * SP) & BM)
	tlo += GAI_SPREAD_LO(z1) + GAI_SPREAD_LO(z2.s) * 2u + GAI_SPREAD_LO(z4.s) * 4u + GAI_SPREAD_LO(z8.s) * 8u + GAI_SPREAD_LO(z16.s) * 16u + GAI_SPREAD_LO(z16.c) * 32u;
	thi += GAI_SPREAD_HI(z1) + GAI_SPREAD_HI(z2.s) * 2u + GAI_SPREAD_HI(z4.s) * 4u + GAI_SPREAD_HI(z8.s) * 8u + GAI_SPREAD_HI(z16.s) * 16u + GAI_SPREAD_HI(z16.c) * 32u;
#undef GAI_SPREAD_LO
#undef GAI_SPREAD_HI
	RowCounts rc;
	rc.plo = tlo * BM;                                   // byte i = rows 0..i   (n <= 96 < 256: no byte overflows)
	rc.phi = thi * BM + (rc.plo >> 24) * BM;
	rc.n = rc.phi >> 24;
	return rc;
}

// step 3: the idx-th move (reference order) -> (from, direction)
// Warp-collective (one REDUX): every lane of the warp must call it; lanes with valid == false do no work.
__device__ __forceinline__ void sweep_select(const RowCounts& rc, u32 idx, bool valid, const Side& my, const Side& en,
                                             const uint16_t* __restrict__ lut, const u64* __restrict__ sqt, int* from, int* dir)
{
	// row = number of rows whose inclusive prefix is <= idx   (SWAR byte compare, all values < 128)
	const u32 y = (idx + 1u) * 0x01010101u;
	const u32 glo = ((rc.plo | 0x80808080u) - y) & 0x80808080u, ghi = ((rc.phi | 0x80808080u) - y) & 0x80808080u;
	const int row = 8 - __popc(glo) - __popc(ghi);
	u32 k = idx;
	if (row > 0) k -= __byte_perm(rc.plo, rc.phi, (u32)((row - 1) & 7)) & 0xffu;
	// own pieces of that row, ascending column
	u32 rowbits = valid ? (byte_of(my.R, (u32)(row & 7)) & 0xffu) : 0u;
	const int maxp = __reduce_max_sync(0xffffffffu, __popc(rowbits));
	int f = 0, d = 0; bool found = false;
	for (int j = 0; j < maxp; ++j) {
		if (rowbits != 0 && !found) {
			const int c = __ffs((int)rowbits) - 1;
			rowbits &= rowbits - 1;
			const int sq = row * 8 + c;
			u32 v = piece_moves_fast(my, en, sq, lut, sqt);
			const u32 cnt = (u32)__popc(v);
			if (k < cnt) {
				for (u32 t = 0; t < k; ++t) v &= v - 1;            // k <= 7
				f = sq; d = __ffs((int)v) - 1; found = true;
			} else k -= cnt;
		}
	}
	*from = f; *dir = d;
}

} // namespace loa
