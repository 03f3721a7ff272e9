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
//              a carry-save adder tree on bit planes (LOP3 only), whose four planes are turned into byte-lane counts by a
//              conflict-free per-lane table (FMA pipe + LSU).  A byte-lane prefix sum gives n and the row that holds move idx.
//   3. piece : inside that row the reference order is by column, then direction: the row's fields are gathered from the
//              line words (one mask per register + IDP.4A), counted per column through the field-count table, and the
//              one cell that holds the move is rank-selected on its 8 move bits.
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
// Multipliers read from the constant bank: ptxas cannot strength-reduce  x * c[..] + y  into LEA / SHF / LOP3 (ALU pipe),
// it stays an IMAD.
__constant__ u32 c_one = 1, c_two = 2, c_four = 4, c_eight = 8, c_14 = 14, c_16 = 16, c_256 = 256, c_64k = 65536;
// dp4a weights (a byte weight in byte position k).  IDP.4A takes its weight from a uniform register; kept together in the
// constant bank they arrive four per LDCU.128 instead of one UMOV each.
struct WeightTab { u32 w37[4], w2[4], w128[4], w8[4]; };
__constant__ WeightTab c_wt = { { 37u, 37u << 8, 37u << 16, 37u << 24 }, { 2u, 2u << 8, 2u << 16, 2u << 24 },
                                { 128u, 128u << 8, 128u << 16, 128u << 24 }, { 8u, 8u << 8, 8u << 16, 8u << 24 } };
#if defined(__CUDA_ARCH__)
#define GAI_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define GAI_WARP_ANY(p) (p)
#endif
#if defined(__CUDA_ARCH__)
typedef u32 LutRef;                                   // shared-memory byte address of the line LUT
typedef u32 TabRef;                                   // shared-memory byte address of a per-lane table, this lane's column
#define GAI_MADC(x, cname, cval, y) ((x) * cname + (y))
#else
typedef const uint16_t* LutRef;                       // host twin (tests/helpers/host_rules.cu): plain pointers
typedef const u32* TabRef;
#define GAI_MADC(x, cname, cval, y) ((x) * (u32)(cval) + (y))
#endif
struct SweepTabs { LutRef lut; TabRef slt, tpc; };
// device: the three tables in the CTA's shared memory (lut: LUT_BYTES; slt, tpc: 256 entries x 32 lanes x 4 B each)
__device__ __forceinline__ SweepTabs sweep_tabs(const uint16_t* lut, const u32* slt, const u32* tpc, u32 lane)
{
	SweepTabs t{};
#if defined(__CUDA_ARCH__)
	t.lut = (u32)__cvta_generic_to_shared(lut);
	t.slt = (u32)__cvta_generic_to_shared(slt) + 4u * lane;
	t.tpc = (u32)__cvta_generic_to_shared(tpc) + 4u * lane;
#endif
	return t;
}
GAI_HDI SweepTabs sweep_tabs_host(const uint16_t* lut, const u32* slt, const u32* tpc)      // host twin: three plain arrays
{
	SweepTabs t{};
#if !defined(__CUDA_ARCH__)
	t.lut = lut; t.slt = slt; t.tpc = tpc;
#endif
	return t;
}
#ifndef GAI_LUT_ADDR_PRMT
#define GAI_LUT_ADDR_PRMT 1
#endif
__constant__ u32 c_wlut = 2u | (255u << 8) | (255u << 16) | (8u << 24);      // 2 o + (255 + 255 + 8) e = 2 (o + 259 e)
static_assert(2 * LUT_STRIDE == 255 + 255 + 8, "the enemy-byte weight of the LUT address is the sum of three dp4a byte weights");
static_assert(2 * LUT_STRIDE == 14 * 37, "the enemy-byte weight of the LUT address is 37 (one dp4a byte) times 14");
template <int I> GAI_HDI u32 half_of(u64 x) { return I < 4 ? (u32)x : (u32)(x >> 32); }
template <int I> GAI_HDI u32 lut16(LutRef lut, u64 my, u64 en)          // LUT word of byte I of a rotation
{
	const u32 m = half_of<I>(my), e = half_of<I>(en);
#if defined(__CUDA_ARCH__)
#if GAI_LUT_ADDR_PRMT
	// v18: one PRMT puts (own byte, enemy byte x 3) into a register, ONE IDP.4A with the byte weights (2, 255, 255, 8)
	// gives 2 o + 518 e + base: three instructions per lookup (PRMT, IDP.4A, LDS) instead of four
	constexpr u32 SEL = (u32)(I & 3) | ((u32)(4 + (I & 3)) * 0x1110u);
	const u32 a = hd_dp4a(hd_byte_perm(m, e, SEL), c_wlut, lut);
#else
	const u32 a = hd_dp4a(e, c_wt.w37[I & 3], 0u) * c_14 + hd_dp4a(m, c_wt.w2[I & 3], lut);
#endif
	u32 v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
#else
	return lut[((m >> (8 * (I & 3))) & 0xffu) + ((e >> (8 * (I & 3))) & 0xffu) * (u32)LUT_STRIDE];
#endif
}
// Short line in the LOW nibble of byte I (both bytes pre-masked to the nibble, phantoms OR-ed in): index o4 + 16 e4.
// The per-lane table has a 128-byte entry pitch (32 lanes x 4 B), so the byte address is 128 o4 + 2048 e4 + lane column.
template <int I> GAI_HDI u32 slt_low(TabRef slt, u64 my, u64 en)
{
	const u32 m = half_of<I>(my), e = half_of<I>(en);
#if defined(__CUDA_ARCH__)
	const u32 a = hd_dp4a(e, c_wt.w128[I & 3], 0u) * c_16 + hd_dp4a(m, c_wt.w128[I & 3], slt);
	u32 v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
#else
	return slt[((m >> (8 * (I & 3))) & 0xffu) + 16u * ((e >> (8 * (I & 3))) & 0xffu)];
#endif
}
// Short line in the HIGH nibble (bytes pre-masked to it): o = 16 o4, e = 16 e4 -> 128 o4 + 2048 e4 = 8 o + 128 e: two IDP.4A
template <int I> GAI_HDI u32 slt_high(TabRef slt, u64 my, u64 en)
{
	const u32 m = half_of<I>(my), e = half_of<I>(en);
#if defined(__CUDA_ARCH__)
	const u32 a = hd_dp4a(e, c_wt.w128[I & 3], hd_dp4a(m, c_wt.w8[I & 3], slt));
	u32 v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
#else
	return slt[(((m >> (8 * (I & 3))) & 0xffu) >> 4) + ((e >> (8 * (I & 3))) & 0xffu)];
#endif
}
// field counts of byte J of a bit-plane register, in byte lanes (tpc_entry)
template <int J> GAI_HDI u32 tpc_of(TabRef tpc, u32 plane)
{
#if defined(__CUDA_ARCH__)
	const u32 a = hd_dp4a(plane, c_wt.w128[J], tpc);
	u32 v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
#else
	return tpc[(plane >> (8 * J)) & 0xffu];
#endif
}
GAI_HDI u32 pack16(u32 lo, u32 hi) { return GAI_MADC(hi, c_64k, 65536, lo); }      // lo | hi << 16 (both < 65536) as one IMAD

// ---- which cells of a diagonal byte belong to which line (compile time).
// D1 byte K: line a = diagonal r - c = K, rows K..7; line b = diagonal r - c = K - 8, rows 0..K-1.
// D2 byte K: line a = diagonal r + c = K, rows 0..K; line b = diagonal r + c = K + 8, rows K+1..7.
// A line of 5+ cells goes to the big LUT with the other line's cells as phantoms (o = e = 1); a line of 2-4 cells sits in
// one nibble and goes to the short-line table with the rest of ITS nibble as phantoms; a one-cell line has no moves.
struct DiagConsts { u64 longA, longB, keepLo, phLo, keepHi, phHi; };
__host__ __device__ constexpr u32 d_first(bool d1, int K, bool a) { return d1 ? (a ? K : 0) : (a ? 0 : K + 1); }          // first row of the line
__host__ __device__ constexpr u32 d_len(bool d1, int K, bool a) { return d1 ? (a ? 8 - K : K) : (a ? K + 1 : 7 - K); }
__host__ __device__ constexpr u32 d_cells(bool d1, int K, bool a) { return ((1u << d_len(d1, K, a)) - 1u) << d_first(d1, K, a); }
__host__ __device__ constexpr DiagConsts make_diag_consts(bool d1)
{
	DiagConsts c{0, 0, 0, 0, 0, 0};
	for (int K = 0; K < 8; ++K) {
		for (int ab = 0; ab < 2; ++ab) {
			const bool a = ab == 0;
			const u32 len = d_len(d1, K, a), cells = d_cells(d1, K, a);
			if (len >= 5) { (a ? c.longA : c.longB) |= (u64)(0xffu & ~cells) << (8 * K); }
			else if (len >= 2) {
				const bool low = d_first(d1, K, a) == 0;                 // short lines start at row 0 or end at row 7
				if (low) { c.keepLo |= (u64)0x0fu << (8 * K); c.phLo |= (u64)(0x0fu & ~cells) << (8 * K); }
				else { c.keepHi |= (u64)0xf0u << (8 * K); c.phHi |= (u64)(0xf0u & ~cells) << (8 * K); }
			}
		}
	}
	return c;
}
template <bool D1> struct DiagIn {                        // the pre-masked copies of one rotation (both colours)
	u64 mA, eA, mB, eB, mLo, eLo, mHi, eHi;
	GAI_HDI DiagIn(u64 my, u64 en)
	{
		constexpr DiagConsts c = make_diag_consts(D1);
		mA = my | c.longA; eA = en | c.longA; mB = my | c.longB; eB = en | c.longB;
		mLo = (my & c.keepLo) | c.phLo; eLo = (en & c.keepLo) | c.phLo;
		mHi = (my & c.keepHi) | c.phHi; eHi = (en & c.keepHi) | c.phHi;
	}
};
// the 16-bit move word of diagonal byte K (both lines; their cells are disjoint, so the parts ADD)
template <bool D1, int K> GAI_HDI u32 sweep_diag(const DiagIn<D1>& d, const SweepTabs& t)
{
	u32 v = 0;
	bool have = false;
#define GAI_ACC(x, cname, cval) do { const u32 x_ = (x); v = have ? GAI_MADC(x_, cname, cval, v) : ((cval) == 1 ? x_ : GAI_MADC(x_, cname, cval, 0u)); have = true; } while (0)
	if (d_len(D1, K, true) >= 5) GAI_ACC(lut16<K>(t.lut, d.mA, d.eA), c_one, 1);
	if (d_len(D1, K, false) >= 5) GAI_ACC(lut16<K>(t.lut, d.mB, d.eB), c_one, 1);
	for (int ab = 0; ab < 2; ++ab) {
		const bool a = ab == 0;
		const u32 len = d_len(D1, K, a);
		if (len < 2 || len >= 5) continue;
		if (d_first(D1, K, a) == 0) GAI_ACC(slt_low<K>(t.slt, d.mLo, d.eLo), c_one, 1);
		else GAI_ACC(slt_high<K>(t.slt, d.mHi, d.eHi), c_256, 256);       // nibble cells 0-3 are rows 4-7
	}
#undef GAI_ACC
	return v;
}
template <int K> GAI_HDI u32 sweep_r(const Side& my, const Side& en, const SweepTabs& t) { return lut16<K>(t.lut, my.R, en.R); }
template <int K> GAI_HDI u32 sweep_c(const Side& my, const Side& en, const SweepTabs& t) { return lut16<K>(t.lut, my.C, en.C); }
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
GAI_HDI RowCounts sweep_rows(const Side& my, const Side& en, const SweepTabs& lut)
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
	// D2 and D1: cell = row as well (word k = (r + c) & 7 / (r - c) & 7)
	const DiagIn<false> dg2(my.D2, en.D2);
	const DiagIn<true> dg1(my.D1, en.D1);
	w[4] = pack16(sweep_diag<false, 0>(dg2, lut), sweep_diag<false, 1>(dg2, lut));
	w[5] = pack16(sweep_diag<false, 2>(dg2, lut), sweep_diag<false, 3>(dg2, lut));
	w[6] = pack16(sweep_diag<false, 4>(dg2, lut), sweep_diag<false, 5>(dg2, lut));
	w[7] = pack16(sweep_diag<false, 6>(dg2, lut), sweep_diag<false, 7>(dg2, lut));
	w[8] = pack16(sweep_diag<true, 0>(dg1, lut), sweep_diag<true, 1>(dg1, lut));
	w[9] = pack16(sweep_diag<true, 2>(dg1, lut), sweep_diag<true, 3>(dg1, lut));
	w[10] = pack16(sweep_diag<true, 4>(dg1, lut), sweep_diag<true, 5>(dg1, lut));
	w[11] = pack16(sweep_diag<true, 6>(dg1, lut), sweep_diag<true, 7>(dg1, lut));

	// positional popcount of the 12 registers: carry-save adders on bit planes
	const FullAdd a0 = fa(w[0], w[1], w[2]), a1 = fa(w[3], w[4], w[5]), a2 = fa(w[6], w[7], w[8]), a3 = fa(w[9], w[10], w[11]);
	const FullAdd b0 = fa(a0.s, a1.s, a2.s);
	const u32 p1 = b0.s ^ a3.s, c5 = b0.s & a3.s;                                    // ones
	const FullAdd t0 = fa(a0.c, a1.c, a2.c), t1 = fa(a3.c, b0.c, c5);
	const u32 p2 = t0.s ^ t1.s, d2 = t0.s & t1.s;                                    // twos
	const FullAdd f4 = fa(t0.c, t1.c, d2);
	const u32 p4 = f4.s, p8 = f4.c;                                                  // fours, eights   (count <= 12)
	// The four planes hold, per (16-bit half, row, sense), one bit of a count <= 12.  Per-row totals: every byte of a plane
	// covers four rows of one half; the field-count table turns it into four byte lanes (popcount of the two sense bits),
	// weighted by the plane and accumulated -- IDP.4A + LDS + IMAD per byte, nothing on the ALU pipe.
	const u32 BM = 0x01010101u;
	tlo = tlo + tpc_of<0>(lut.tpc, p1) + tpc_of<2>(lut.tpc, p1);          // weight 1: one three-input add each (the kernel is issue bound)
	thi = thi + tpc_of<1>(lut.tpc, p1) + tpc_of<3>(lut.tpc, p1);
	tlo = GAI_MADC(tpc_of<0>(lut.tpc, p2), c_two, 2, tlo); tlo = GAI_MADC(tpc_of<2>(lut.tpc, p2), c_two, 2, tlo);
	thi = GAI_MADC(tpc_of<1>(lut.tpc, p2), c_two, 2, thi); thi = GAI_MADC(tpc_of<3>(lut.tpc, p2), c_two, 2, thi);
	tlo = GAI_MADC(tpc_of<0>(lut.tpc, p4), c_four, 4, tlo); tlo = GAI_MADC(tpc_of<2>(lut.tpc, p4), c_four, 4, tlo);
	thi = GAI_MADC(tpc_of<1>(lut.tpc, p4), c_four, 4, thi); thi = GAI_MADC(tpc_of<3>(lut.tpc, p4), c_four, 4, thi);
	// the eights plane is set only where EIGHT or more of the twelve words carry the same (row, sense) flag -- three own
	// pieces in one row all moving the same way along three line types -- so it is empty in almost every position: the
	// whole warp skips its four lookups unless some lane needs them (every caller runs this function warp-synchronously)
	if (GAI_WARP_ANY(p8 != 0)) {
		tlo = GAI_MADC(tpc_of<0>(lut.tpc, p8), c_eight, 8, tlo); tlo = GAI_MADC(tpc_of<2>(lut.tpc, p8), c_eight, 8, tlo);
		thi = GAI_MADC(tpc_of<1>(lut.tpc, p8), c_eight, 8, thi); thi = GAI_MADC(tpc_of<3>(lut.tpc, p8), c_eight, 8, thi);
	}
	rc.plo = tlo * BM;                                   // byte i = rows 0..i   (n <= 96 < 256: no byte overflows)
	rc.phi = thi * BM + (rc.plo >> 24) * BM;
	rc.n = rc.phi >> 24;
	return rc;
}

// step 3: the idx-th move (reference order) -> (from, direction).  Straight-line code, no lookups: the moves of
// the selected row are gathered from the line words (one 2-bit field per column and orientation), interleaved into the
// reference order (column, then W E | N S | SW NE | SE NW) and rank-selected.
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
// step 3.  Only ONE cell's byte of move bits is ever needed, so the row is not interleaved into 64 bits: the per-column
// move counts come from the field-count table (8 conflict-free lookups), a byte-lane prefix sum and a SWAR compare give
// the column, the four 2-bit fields of that column are pulled out with two shifts and two IDP.4A, and the rank-select
// runs on 8 bits.
GAI_HDI void sweep_select(const RowCounts& rc, u32 idx, const SweepTabs& t, int* from, int* dir)
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
	// C: field `row` of column word c -> position c  (the field holds (S, N); swapped below, counting does not care)
	const u32 xc = gather_fields<0, 1, 2, 3, 4, 5, 6, 7>(rc.w + 0, M, sh4);
	// D2: byte k = (row + c) & 7  ->  column c = (k - row) & 7 : gather by k, rotate right by `row` fields
	const u32 y2 = gather_fields<0, 1, 2, 3, 4, 5, 6, 7>(rc.w + 4, M, sh4);
	const u32 x2 = rotl16(y2, (int)((16u - sh) & 15u));
	// D1: byte k = (row - c) & 7 -> column c = (row - k) & 7 : gather reversed ((8-k)&7), rotate left by `row` fields
	const u32 y1 = gather_fields<0, 7, 6, 5, 4, 3, 2, 1>(rc.w + 8, M, sh4);
	const u32 x1 = rotl16(y1, (int)sh);
	// moves per column, one byte lane each (<= 8), inclusive prefix sums
	const u32 A = pack16(xr, xc), B = pack16(x1, x2);
	const u32 cl = (tpc_of<0>(t.tpc, A) + tpc_of<2>(t.tpc, A)) + (tpc_of<0>(t.tpc, B) + tpc_of<2>(t.tpc, B));     // three-input adds
	const u32 ch = (tpc_of<1>(t.tpc, A) + tpc_of<3>(t.tpc, A)) + (tpc_of<1>(t.tpc, B) + tpc_of<3>(t.tpc, B));
	const u32 pl = cl * 0x01010101u, ph = ch * 0x01010101u + (pl >> 24) * 0x01010101u;
	const u32 yk = (k + 1u) * 0x01010101u;
	const u32 hlo = ((pl | 0x80808080u) - yk) & 0x80808080u, hhi = ((ph | 0x80808080u) - yk) & 0x80808080u;
	const u32 col = (u32)(8 - hd_popc(hlo) - hd_popc(hhi)) & 7u;
	if (col > 0) k -= hd_byte_perm(pl, ph, col - 1u) & 0xffu;
	// the cell's 8 move bits in reference order: W E | N S | SW NE | SE NW
	const u32 cs = 2u * col;
	const u32 fa = (A >> cs) & 0x00030003u, fb = (B >> cs) & 0x00030003u;       // (R, C) and (D1, D2) fields of the column
	u32 v = hd_dp4a(fb, 0x00400010u, hd_dp4a(fa, 0x00040001u, 0u));              // R | C << 2 | D1 << 4 | D2 << 6
	v = (v & 0xf3u) | ((v & 4u) << 1) | ((v >> 1) & 4u);                         // C field (S, N) -> (N, S)
	int pos = 0;
	{ u32 c = hd_popc(v & 0xfu); bool g = k >= c; k -= g ? c : 0u; pos += g ? 4 : 0; v = g ? v >> 4 : v;
	  c = hd_popc(v & 0x3u); g = k >= c; k -= g ? c : 0u; pos += g ? 2 : 0; v = g ? v >> 2 : v;
	  pos += (k >= (v & 1u)) ? 1 : 0; }
	*from = (int)(row * 8u + col);
	*dir = pos;
}

} // namespace loa
