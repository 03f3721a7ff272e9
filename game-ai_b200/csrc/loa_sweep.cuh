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
//              row, so the per-row totals are a POSITIONAL popcount over 24 words:
//              a carry-save adder tree on bit planes (LOP3 only), folded to 6-bit counts and spread to byte lanes with
//              integer multiplies (FMA pipe).  A byte-lane prefix sum gives n and the row that holds move idx.
//   3. piece : inside that row the reference order is by column, then direction: the row's 8 x 8 move bits are
//              gathered from the line words with shifts and rotations, interleaved, and rank-selected.  No lookups.
// Move ORDER is the reference's (LinesOfActionZasady.cpp:180-223): ascending square, then W E N S SW NE SE NW.
#pragma once
#include "loa_fast.cuh"

namespace loa {

// The ALU pipe (LOP3/SHF/PRMT/SEL/IADD3) is the binding unit of this kernel; IMAD and IDP.4A issue on the other
// (FMA-heavy) pipe at the same rate (profiles/r01_ubench_pipes.json: 0.5 warp-instructions / clock / sub-partition each,
// 0.93 when interleaved).  So the LUT byte address  2 * o + 516 * e + base  is computed without touching the ALU pipe:
// IDP.4A with a one-byte weight extracts AND scales a byte of the board register in a single instruction,
//     t = dp4a(enemy_half, 129 << 8k, 0);  u = dp4a(own_half, 2 << 8k, base);  addr = t * 4 + u
// and the 4 is read from the constant bank so that ptxas cannot turn the last IMAD into a LEA (ALU pipe).
__constant__ u32 c_four = 4, c_64k = 65536;
#if defined(__CUDA_ARCH__)
typedef u32 LutRef;                                   // shared-memory byte address of the line LUT
#else
typedef const uint16_t* LutRef;                       // host twin (tests/helpers/host_rules.cu): a plain pointer
#endif
GAI_HDI LutRef lut_ref(const uint16_t* p)
{
#if defined(__CUDA_ARCH__)
	return (u32)__cvta_generic_to_shared(p);
#else
	return p;
#endif
}
static_assert(LUT_STRIDE % 2 == 0 && LUT_STRIDE / 2 <= 255, "the enemy-byte weight of the LUT address must fit one dp4a byte");
template <int I> GAI_HDI u32 lut16(LutRef lut, u64 my, u64 en)          // LUT word of byte I of a rotation
{
	const u32 m = I < 4 ? (u32)my : (u32)(my >> 32), e = I < 4 ? (u32)en : (u32)(en >> 32);
#if defined(__CUDA_ARCH__)
	const u32 a = hd_dp4a(e, (u32)(LUT_STRIDE / 2) << (8 * (I & 3)), 0u) * c_four + hd_dp4a(m, 2u << (8 * (I & 3)), lut);
	u32 v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
#else
	return lut[((m >> (8 * (I & 3))) & 0xffu) + ((e >> (8 * (I & 3))) & 0xffu) * (u32)LUT_STRIDE];
#endif
}
GAI_HDI u32 pack16(u32 lo, u32 hi)                      // lo | hi << 16 (both < 65536) as one IMAD
{
#if defined(__CUDA_ARCH__)
	return hi * c_64k + lo;
#else
	return lo | (hi << 16);
#endif
}

// phantom masks of the two diagonals sharing byte K (compile time).  D1 byte K: the diagonal r - c = K holds rows K..7
// (line a), the diagonal r - c = K - 8 rows 0..K-1 (line b).  D2 byte K: r + c = K holds rows 0..K, r + c = K + 8 rows K+1..7.
template <int K> struct D1Ph { static constexpr u32 a = (1u << K) - 1u, b = (0xffu << K) & 0xffu; static constexpr int len_a = 8 - K, len_b = K; };
template <int K> struct D2Ph { static constexpr u32 a = (0xffu << (K + 1)) & 0xffu, b = (1u << (K + 1)) - 1u; static constexpr int len_a = K + 1, len_b = 7 - K; };

template <int K> GAI_HDI u32 sweep_r(const Side& my, const Side& en, LutRef lut) { return lut16<K>(lut, my.R, en.R); }
template <int K> GAI_HDI u32 sweep_c(const Side& my, const Side& en, LutRef lut) { return lut16<K>(lut, my.C, en.C); }
// The phantom masks are OR-ed into whole registers (4 bytes at a time) before the byte extraction: myA = board | PHA.
template <template <int> class PH> struct PhWords {
	static constexpr u64 A = (u64)PH<0>::a | ((u64)PH<1>::a << 8) | ((u64)PH<2>::a << 16) | ((u64)PH<3>::a << 24) | ((u64)PH<4>::a << 32) | ((u64)PH<5>::a << 40) | ((u64)PH<6>::a << 48) | ((u64)PH<7>::a << 56);
	static constexpr u64 B = (u64)PH<0>::b | ((u64)PH<1>::b << 8) | ((u64)PH<2>::b << 16) | ((u64)PH<3>::b << 24) | ((u64)PH<4>::b << 32) | ((u64)PH<5>::b << 40) | ((u64)PH<6>::b << 48) | ((u64)PH<7>::b << 56);
};
template <int K, class PH> GAI_HDI u32 sweep_diag(u64 myA, u64 enA, u64 myB, u64 enB, LutRef lut)
{
	u32 v = 0;
	if (PH::len_a >= 2) v = lut16<K>(lut, myA, enA);       // a one-cell line has no moves
	if (PH::len_b >= 2) v |= lut16<K>(lut, myB, enB);      // the two lines own disjoint cells
	return v;
}
GAI_HDI u32 rotl16(u32 x, int s)                     // x < 65536, 0 <= s < 16
{
	const u32 d = x * 0x10001u;
	return (d >> ((16 - s) & 15)) & 0xffffu;
}

struct FullAdd { u32 s, c; };
GAI_HDI FullAdd fa(u32 a, u32 b, u32 c) { FullAdd r; r.s = a ^ b ^ c; r.c = (a & b) | (c & (a ^ b)); return r; }

struct RowCounts {
	u32 plo, phi;      // inclusive prefix sums of the per-row move totals, one byte per row (rows 0-3 / 4-7)
	u32 n;             // total number of legal moves
	u32 rp[4];         // R words, two per register (rows 2j | 2j+1 << 16), fields by column
	u32 w[12];         // C (w0-3), D2 (w4-7), D1 (w8-11) words, two per register, fields by ROW
};

// steps 1 + 2.  Every lane executes exactly this straight-line code.
GAI_HDI RowCounts sweep_rows(const Side& my, const Side& en, LutRef lut)
{
	// ---- R lines: all cells of a word lie in one row
	const u32 r0 = sweep_r<0>(my, en, lut), r1 = sweep_r<1>(my, en, lut), r2 = sweep_r<2>(my, en, lut), r3 = sweep_r<3>(my, en, lut);
	const u32 r4 = sweep_r<4>(my, en, lut), r5 = sweep_r<5>(my, en, lut), r6 = sweep_r<6>(my, en, lut), r7 = sweep_r<7>(my, en, lut);
	u32 tlo = hd_popc(r0) + (hd_popc(r1) << 8) + (hd_popc(r2) << 16) + (hd_popc(r3) << 24);
	u32 thi = hd_popc(r4) + (hd_popc(r5) << 8) + (hd_popc(r6) << 16) + (hd_popc(r7) << 24);

	// ---- C, D1, D2 lines: one cell per row; two 16-bit words per register
	RowCounts rc;
	u32* w = rc.w;
	rc.rp[0] = pack16(r0, r1); rc.rp[1] = pack16(r2, r3); rc.rp[2] = pack16(r4, r5); rc.rp[3] = pack16(r6, r7);
	w[0] = pack16(sweep_c<0>(my, en, lut), sweep_c<1>(my, en, lut));
	w[1] = pack16(sweep_c<2>(my, en, lut), sweep_c<3>(my, en, lut));
	w[2] = pack16(sweep_c<4>(my, en, lut), sweep_c<5>(my, en, lut));
	w[3] = pack16(sweep_c<6>(my, en, lut), sweep_c<7>(my, en, lut));
	// D2: cell = row
	const u64 m2a = my.D2 | PhWords<D2Ph>::A, e2a = en.D2 | PhWords<D2Ph>::A, m2b = my.D2 | PhWords<D2Ph>::B, e2b = en.D2 | PhWords<D2Ph>::B;
	const u64 m1a = my.D1 | PhWords<D1Ph>::A, e1a = en.D1 | PhWords<D1Ph>::A, m1b = my.D1 | PhWords<D1Ph>::B, e1b = en.D1 | PhWords<D1Ph>::B;
	w[4] = pack16(sweep_diag<0, D2Ph<0>>(m2a, e2a, m2b, e2b, lut), sweep_diag<1, D2Ph<1>>(m2a, e2a, m2b, e2b, lut));
	w[5] = pack16(sweep_diag<2, D2Ph<2>>(m2a, e2a, m2b, e2b, lut), sweep_diag<3, D2Ph<3>>(m2a, e2a, m2b, e2b, lut));
	w[6] = pack16(sweep_diag<4, D2Ph<4>>(m2a, e2a, m2b, e2b, lut), sweep_diag<5, D2Ph<5>>(m2a, e2a, m2b, e2b, lut));
	w[7] = pack16(sweep_diag<6, D2Ph<6>>(m2a, e2a, m2b, e2b, lut), sweep_diag<7, D2Ph<7>>(m2a, e2a, m2b, e2b, lut));
	// D1: cell = row as well (word k = (r - c) & 7)
	w[8] = pack16(sweep_diag<0, D1Ph<0>>(m1a, e1a, m1b, e1b, lut), sweep_diag<1, D1Ph<1>>(m1a, e1a, m1b, e1b, lut));
	w[9] = pack16(sweep_diag<2, D1Ph<2>>(m1a, e1a, m1b, e1b, lut), sweep_diag<3, D1Ph<3>>(m1a, e1a, m1b, e1b, lut));
	w[10] = pack16(sweep_diag<4, D1Ph<4>>(m1a, e1a, m1b, e1b, lut), sweep_diag<5, D1Ph<5>>(m1a, e1a, m1b, e1b, lut));
	w[11] = pack16(sweep_diag<6, D1Ph<6>>(m1a, e1a, m1b, e1b, lut), sweep_diag<7, D1Ph<7>>(m1a, e1a, m1b, e1b, lut));

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
	rc.plo = tlo * BM;                                   // byte i = rows 0..i   (n <= 96 < 256: no byte overflows)
	rc.phi = thi * BM + (rc.plo >> 24) * BM;
	rc.n = rc.phi >> 24;
	return rc;
}

// step 3: the idx-th move (reference order) -> (from, direction).  Straight-line code, no lookups: the moves of
// the selected row are gathered from the line words (one 2-bit field per column and orientation), interleaved into the
// reference order (column, then W E | N S | SW NE | SE NW) and rank-selected.
GAI_HDI u32 spread2to4(u32 x)          // 8 two-bit fields (16 bits) -> 8 nibbles
{
	x = (x | (x << 8)) & 0x00ff00ffu;
	x = (x | (x << 4)) & 0x0f0f0f0fu;
	x = (x | (x << 2)) & 0x33333333u;
	return x;
}
GAI_HDI u32 spread4to8(u32 x)          // 4 nibbles (16 bits) -> 4 bytes
{
	x = (x | (x << 8)) & 0x00ff00ffu;
	x = (x | (x << 4)) & 0x0f0f0f0fu;
	return x;
}
// Field `row` of each of the 8 words packed two per register in p[0..3] -> one 16-bit word, word j's field at 2-bit
// position Pj.  M = 0x00030003 << 2*row isolates the field (one LOP3 per register); after that only ONE byte of each
// 16-bit half is non-zero, so a dp4a with the weight 4^P in both bytes of the half moves the field to position P whatever
// the row is: the sums come out scaled by 4^(row & 3) and are shifted back once at the end.  Positions 0-3 and 4-7 are
// collected separately (a dp4a weight is one byte).  4 LOP3 + 1 SHF on the ALU pipe, the rest IDP.4A / IMAD.
__host__ __device__ constexpr u32 gf_w(int P, bool hi) { return (hi ? (P >= 4) : (P < 4)) ? (1u << (2 * (P & 3))) * 0x0101u : 0u; }
template <int PA, int PB, bool HI> GAI_HDI u32 gf_acc(u32 m, u32 acc)
{
	constexpr u32 W = gf_w(PA, HI) | (gf_w(PB, HI) << 16);
	return W ? hd_dp4a(m, W, acc) : acc;
}
template <int P0, int P1, int P2, int P3, int P4, int P5, int P6, int P7>
GAI_HDI u32 gather_fields(const u32* p, u32 M, u32 sh4)
{
	const u32 m0 = p[0] & M, m1 = p[1] & M, m2 = p[2] & M, m3 = p[3] & M;
	const u32 lo = gf_acc<P6, P7, false>(m3, gf_acc<P4, P5, false>(m2, gf_acc<P2, P3, false>(m1, gf_acc<P0, P1, false>(m0, 0u))));
	const u32 hi = gf_acc<P6, P7, true>(m3, gf_acc<P4, P5, true>(m2, gf_acc<P2, P3, true>(m1, gf_acc<P0, P1, true>(m0, 0u))));
	return (hi * 256u + lo) >> sh4;
}
GAI_HDI void sweep_select(const RowCounts& rc, u32 idx, int* from, int* dir)
{
	// row = number of rows whose inclusive prefix is <= idx   (SWAR byte compare, all values < 128)
	const u32 y = (idx + 1u) * 0x01010101u;
	const u32 glo = ((rc.plo | 0x80808080u) - y) & 0x80808080u, ghi = ((rc.phi | 0x80808080u) - y) & 0x80808080u;
	const u32 row = (u32)(8 - hd_popc(glo) - hd_popc(ghi)) & 7u;
	u32 k = idx;
	if (row > 0) k -= hd_byte_perm(rc.plo, rc.phi, row - 1u) & 0xffu;
	const u32 sh = 2u * row, sh4 = sh & 6u, M = 0x00030003u << sh;
	// R: the row's own word (fields by column: W, E)
	const u32 rsel = (row & 4u) ? ((row & 2u) ? rc.rp[3] : rc.rp[2]) : ((row & 2u) ? rc.rp[1] : rc.rp[0]);
	const u32 xr = (rsel >> (16u * (row & 1u))) & 0xffffu;
	// C: field `row` of column word c -> position c; the field holds (S, N), the reference wants (N, S)
	const u32 xc0 = gather_fields<0, 1, 2, 3, 4, 5, 6, 7>(rc.w + 0, M, sh4);
	const u32 xc = ((xc0 & 0x5555u) << 1) | ((xc0 >> 1) & 0x5555u);
	// D2: byte k = (row + c) & 7  ->  column c = (k - row) & 7 : gather by k, rotate right by `row` fields
	const u32 y2 = gather_fields<0, 1, 2, 3, 4, 5, 6, 7>(rc.w + 4, M, sh4);
	const u32 x2 = rotl16(y2, (int)((16u - sh) & 15u));
	// D1: byte k = (row - c) & 7 -> column c = (row - k) & 7 : gather reversed ((8-k)&7), rotate left by `row` fields
	const u32 y1 = gather_fields<0, 7, 6, 5, 4, 3, 2, 1>(rc.w + 8, M, sh4);
	const u32 x1 = rotl16(y1, (int)sh);
	// interleave: byte c of z = xr[c] | xc[c] << 2 | x1[c] << 4 | x2[c] << 6
	const u32 a = spread2to4(xr) | (spread2to4(xc) << 2);          // nibble c = R,C fields
	const u32 b = spread2to4(x1) | (spread2to4(x2) << 2);          // nibble c = D1,D2 fields
	const u32 zlo = spread4to8(a & 0xffffu) | (spread4to8(b & 0xffffu) << 4);
	const u32 zhi = spread4to8(a >> 16) | (spread4to8(b >> 16) << 4);
	// k-th set bit of (zhi:zlo), branch free
	const u32 clo = hd_popc(zlo);
	const bool hi = k >= clo;
	u32 x = hi ? zhi : zlo; k -= hi ? clo : 0u;
	int pos = hi ? 32 : 0;
	{ u32 c = hd_popc(x & 0xffffu); bool g = k >= c; k -= g ? c : 0u; pos += g ? 16 : 0; x = g ? x >> 16 : x;
	  c = hd_popc(x & 0xffu); g = k >= c; k -= g ? c : 0u; pos += g ? 8 : 0; x = g ? x >> 8 : x;
	  c = hd_popc(x & 0xfu); g = k >= c; k -= g ? c : 0u; pos += g ? 4 : 0; x = g ? x >> 4 : x;
	  c = hd_popc(x & 0x3u); g = k >= c; k -= g ? c : 0u; pos += g ? 2 : 0; x = g ? x >> 2 : x;
	  pos += (k >= (x & 1u)) ? 1 : 0; }
	*from = (int)(row * 8u) + (pos >> 3);
	*dir = pos & 7;
}

} // namespace loa
