// game-ai_b200/csrc/loa_euler.cuh -- Euler characteristic of a colour's 8-connected figure, maintained move by move.
//
// calcIfTerminal (LinesOfActionZasady.cpp:52-58) asks "is this colour ONE 8-connected group?".  The flood fill that answers
// it (loa_rules.cuh connected) ran in 46 % of warp-plies for a single lane of 32 (ncu r01_v15: 15 % of all issued
// instructions went into connectivity).  An exact necessary condition that costs a table lookup per changed square:
//     chi = components - holes                      (8-connected foreground, 4-connected background)
// so one group means chi = 1 - holes <= 1, and chi > 1 proves >= 2 components: only chi <= 1 needs the fill (a single
// group, or the rare k groups with at least k - 1 holes).  chi is a sum of local terms (vertices - edges +
// triangles - tetrahedra of the 8-adjacency clique complex), so adding or removing the piece on square p changes it by a
// function of p's eight neighbours alone: delta_add[neighbourhood], 384 bytes.
//
// Neighbourhood index of square p = 8 r + c in board b: the three row bytes r-1, r, r+1 picked with one PRMT (selector from
// the per-square table), masked to columns c-1..c+1 without the centre (rows / columns off the board: mask 0), rotated right
// by c-1 (c = 0: left by one) so that the 3 x 3 window sits at bits 0-2 / 8,10 / 16-18, then one IDP.4A with the byte weights
// (1, 64, 8) -> 0..383.
#pragma once
#include "loa_rules.cuh"

namespace loa {

// per square: lo = column/row premask of the three picked bytes (24 bits); hi = PRMT selector (bits 0-15) | rotate amount << 16
struct NbxTable { u64 w[64]; };
__host__ __device__ constexpr NbxTable make_nbx_table()
{
	NbxTable t{};
	for (int p = 0; p < 64; ++p) {
		const int r = p >> 3, c = p & 7;
		u32 colbits = 0;
		for (int d = -1; d <= 1; ++d) if (c + d >= 0 && c + d < 8) colbits |= 1u << (c + d);
		u32 mask = 0, sel = 0;
		for (int j = 0; j < 3; ++j) {
			const int row = r - 1 + j;
			if (row >= 0 && row < 8) { mask |= (j == 1 ? (colbits & ~(1u << c)) : colbits) << (8 * j); sel |= (u32)row << (4 * j); }
		}
		const u32 rot = (u32)((c + 31) & 31);              // rotate right by c - 1; c = 0: right by 31 = left by one
		t.w[p] = (u64)mask | ((u64)(sel | (rot << 16)) << 32);
	}
	return t;
}

// delta of chi when the centre of a 3 x 3 window is ADDED; the window's other cells as index bits: 0-2 = row below (W, C, E
// columns), 6-8 -> (idx >> 6): bit 0 = west, bit 2 = east of the centre row, 3-5 = row above  (index = b0 + 8 b2 + 64 b1)
__host__ __device__ constexpr int chi_delta_add(u32 idx)
{
	const u32 below = idx & 7u, above = (idx >> 3) & 7u, mid = (idx >> 6) & 7u;
	bool g[3][3] = { { false, false, false }, { false, false, false }, { false, false, false } };
	for (int c = 0; c < 3; ++c) { g[0][c] = (below >> c) & 1u; g[2][c] = (above >> c) & 1u; }
	g[1][0] = mid & 1u; g[1][2] = (mid >> 2) & 1u;
	// simplices that contain the centre (1,1): edges to each set neighbour, triangles = pairs of set neighbours that are
	// 8-adjacent to each other, tetrahedra = 2 x 2 blocks through the centre with all three other cells set
	int e = 0, t = 0, q = 0;
	for (int a = 0; a < 9; ++a) {
		const int ar = a / 3, ac = a % 3;
		if (a == 4 || !g[ar][ac]) continue;
		++e;
		for (int b = a + 1; b < 9; ++b) {
			const int br = b / 3, bc = b % 3;
			if (b == 4 || !g[br][bc]) continue;
			const int dr = ar > br ? ar - br : br - ar, dc = ac > bc ? ac - bc : bc - ac;
			if (dr <= 1 && dc <= 1) ++t;
		}
	}
	for (int br = 0; br < 2; ++br) for (int bc = 0; bc < 2; ++bc) {           // the four 2 x 2 blocks around the centre
		bool full = true;
		for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) if (!(br + i == 1 && bc + j == 1) && !g[br + i][bc + j]) full = false;
		if (full) ++q;
	}
	return 1 - e + t - q;
}
constexpr int CHI_TAB_SIZE = 384;
struct ChiTable { signed char d[CHI_TAB_SIZE]; };
__host__ __device__ constexpr ChiTable make_chi_table()
{
	ChiTable t{};
	for (int i = 0; i < CHI_TAB_SIZE; ++i) t.d[i] = (signed char)chi_delta_add((u32)i);
	return t;
}

GAI_HDI u32 hd_rotr32(u32 x, u32 s)
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(x, x, s);
#else
	s &= 31u; return s ? (x >> s) | (x << (32u - s)) : x;
#endif
}
// index of the 3 x 3 window around p in board b (the centre itself is masked out)
GAI_HDI u32 chi_index(u64 b, u64 nbx_p)
{
	const u32 ctl = (u32)(nbx_p >> 32);
	const u32 x = hd_byte_perm((u32)b, (u32)(b >> 32), ctl) & (u32)nbx_p;
	const u32 y = hd_rotr32(x, ctl >> 16);
	return hd_dp4a(y, 0x00084001u, 0u);
}
// chi of a whole board, piece by piece (leaf preparation; not on the per-ply path)
GAI_HDI int chi_of(u64 b, const u64* __restrict__ nbx, const signed char* __restrict__ tab)
{
	int chi = 0; u64 acc = 0;
	for (u64 rest = b; rest; rest &= rest - 1) {
		const int p = hd_ffsll(rest) - 1;
		chi += tab[chi_index(acc, nbx[p])];
		acc |= 1ull << p;
	}
	return chi;
}

} // namespace loa
