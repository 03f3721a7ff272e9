// game-ai_b200/csrc/loa_rules.cuh -- Lines-of-Action rules as sm_100a device code.
//
// Device restatement of c++/LinesOfActionZasady/LinesOfActionZasady.cpp (reference, path:line cited per
// function).  Positions are two 64-bit bitboards held in registers (bit = row*8+col, bit0 = A1, :43).
// This header is the BASELINE formulation ("v1"): one piece at a time, 64-bit line masks, used by the
// tier-1 sweep kernels and as the in-kernel cross-check for the faster rollout formulations.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define GAI_HDI __host__ __device__ __forceinline__

namespace loa {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u64 FILE_A = 0x0101010101010101ull;
constexpr u64 FILE_H = 0x8080808080808080ull;
constexpr u64 OPEN_BLACK = 0x7e0000000000007eull;   // LinesOfActionZasady.cpp:362
constexpr u64 OPEN_WHITE = 0x0081818181818100ull;   // :363

enum { OUT_WHITE = 0, OUT_BLACK = 1, OUT_DRAW = 2, OUT_CAPPED = 3 };

// ---- intrinsics with a host twin: the formulations built on them (loa_sweep.cuh) also compile for the CPU so that
// `pytest -m "not gpu"` (tests/helpers/host_rules.cu) checks their logic against the oracle before any GPU time is spent.
GAI_HDI u32 hd_popc(u32 x)
{
#if defined(__CUDA_ARCH__)
	return (u32)__popc(x);
#else
	return (u32)__builtin_popcount(x);
#endif
}
GAI_HDI u32 hd_popcll(u64 x)
{
#if defined(__CUDA_ARCH__)
	return (u32)__popcll(x);
#else
	return (u32)__builtin_popcountll(x);
#endif
}
GAI_HDI int hd_ffsll(u64 x)
{
#if defined(__CUDA_ARCH__)
	return __ffsll((long long)x);
#else
	return __builtin_ffsll((long long)x);
#endif
}
GAI_HDI u32 hd_byte_perm(u32 a, u32 b, u32 sel)           // PRMT, default mode (selector nibbles 0-7, no sign replication used here)
{
#if defined(__CUDA_ARCH__)
	u32 r;                                          // prmt.b32 directly: __byte_perm() first masks the selector with 0x7777 (one more LOP3);
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));      // every selector used here has nibbles 0-7 (no sign replication)
	return r;
#else
	const u64 v = (u64)a | ((u64)b << 32);
	u32 r = 0;
	for (int i = 0; i < 4; ++i) r |= (u32)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
	return r;
#endif
}
GAI_HDI u32 hd_dp4a(u32 a, u32 b, u32 c)                  // IDP.4A.U8.U8: c + sum of the four byte products
{
#if defined(__CUDA_ARCH__)
	return __dp4a(a, b, c);
#else
	for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
	return c;
#endif
}

// ---- line masks (LinesOfActionZasady.cpp:82-135) --------------------------------------------------
__host__ __device__ constexpr u64 row_mask(int p) { return 0xFFull << (p & 56); }
__host__ __device__ constexpr u64 col_mask(int p) { return FILE_A << (p & 7); }
// "left" diagonal A1->H8 (row-col const, :93-113); "right" diagonal H1->A8 (row+col const, :114-135)
__host__ __device__ constexpr u64 ldiag_mask_slow(int p)
{
	u64 m = 0; const int d = (p >> 3) - (p & 7);
	for (int c = 0; c < 8; ++c) { const int r = c + d; if (r >= 0 && r < 8) m |= 1ull << (r * 8 + c); }
	return m;
}
__host__ __device__ constexpr u64 rdiag_mask_slow(int p)
{
	u64 m = 0; const int s = (p >> 3) + (p & 7);
	for (int c = 0; c < 8; ++c) { const int r = s - c; if (r >= 0 && r < 8) m |= 1ull << (r * 8 + c); }
	return m;
}
struct DiagTables { u64 ld[64]; u64 rd[64]; };
__host__ __device__ constexpr DiagTables make_diag_tables()
{
	DiagTables t{};
	for (int p = 0; p < 64; ++p) { t.ld[p] = ldiag_mask_slow(p); t.rd[p] = rdiag_mask_slow(p); }
	return t;
}

// ---- connectivity (LinesOfActionZasady.cpp:52-81, 271-305) ----------------------------------------
// 8-neighbourhood dilation of a set of squares (centre included).
GAI_HDI u64 dilate8(u64 g)
{
	const u64 h = g | ((g << 1) & ~FILE_A) | ((g >> 1) & ~FILE_H);
	return h | (h << 8) | (h >> 8);
}
// calcIfTerminal(board) (:52-58): exactly one piece, or exactly one 8-connected group.
// Flood fill by shift-and-mask dilation from the lowest piece until the fixpoint.
GAI_HDI bool connected(u64 b)
{
	if (b == 0) return false;                 // 0 groups, popcount 0
	u64 g = b & (0 - b);
	if (g == b) return true;                  // one piece
	for (;;) {
		const u64 n = dilate8(g) & b;
		if (n == g) break;
		g = n;
	}
	return g == b;
}
// Cheap necessary condition used before the flood fill: a colour with >= 2 pieces cannot be connected
// if some piece has no neighbour of its own colour.
GAI_HDI bool has_isolated_piece(u64 b)
{
	const u64 h = ((b << 1) & ~FILE_A) | ((b >> 1) & ~FILE_H);
	const u64 v = h | b;
	const u64 nbr = h | (v << 8) | (v >> 8);
	return (b & ~nbr) != 0;
}
GAI_HDI bool connected_fast(u64 b)
{
	if (b == 0) return false;
	if ((b & (b - 1)) == 0) return true;
	if (has_isolated_piece(b)) return false;
	return connected(b);
}

// ---- move generation (LinesOfActionZasady.cpp:136-155, 167-232) -----------------------------------
// Direction order of the reference (:194-223): 0 W(-1) 1 E(+1) 2 N(+8) 3 S(-8) 4 SW(-9) 5 NE(+9) 6 SE(-7) 7 NW(+7)
GAI_HDI int dir_step(int d)
{
	// signed bytes -1,+1,+8,-8,-9,+9,-7,+7 packed little-endian (no local-memory table)
	return (int)(signed char)(0x07F909F7F80801FFull >> (8 * d));
}

// validity of the two senses of one line for the piece on `pos`.  Returns bit0 = towards higher squares,
// bit1 = towards lower squares; *cnt = pieces of both colours on the line (:191-192).
__device__ __forceinline__ u32 line_moves(u64 my, u64 en, u64 L, int step, int pos, int* cnt)
{
	const int n = __popcll((my | en) & L);
	const u64 eL = en & L;          // only ENEMY pieces block (:147-148); own pieces may be jumped
	const u64 okT = L & ~my;        // target must be on the line and not an own piece (:142)
	u32 v = 0;
	const int tu = pos + n * step;
	if (tu < 64 && ((okT >> tu) & 1ull) && (eL & ((1ull << tu) - (2ull << pos))) == 0) v |= 1u;
	const int td = pos - n * step;
	if (td >= 0 && ((okT >> td) & 1ull) && (eL & ((1ull << pos) - (2ull << td))) == 0) v |= 2u;
	*cnt = n;
	return v;
}

// 8-bit move mask of the piece on `pos`, bit d = direction d (reference order) is legal.
__device__ __forceinline__ u32 piece_moves(u64 my, u64 en, int pos, const u64* __restrict__ ld, const u64* __restrict__ rd)
{
	int n;
	const u32 r = line_moves(my, en, row_mask(pos), 1, pos, &n);   // bit0 = E, bit1 = W
	const u32 c = line_moves(my, en, col_mask(pos), 8, pos, &n);   // bit0 = N, bit1 = S
	const u32 l = line_moves(my, en, ld[pos], 9, pos, &n);         // bit0 = NE, bit1 = SW
	const u32 q = line_moves(my, en, rd[pos], 7, pos, &n);         // bit0 = NW, bit1 = SE
	return ((r >> 1) & 1u) | ((r & 1u) << 1) | ((c & 1u) << 2) | ((c >> 1) << 3)
	     | ((l >> 1) << 4) | ((l & 1u) << 5) | ((q >> 1) << 6) | ((q & 1u) << 7);
}

// target square of the move of the piece on `pos` in direction d
GAI_HDI int move_target(u64 all, int pos, int d, const u64* __restrict__ ld, const u64* __restrict__ rd)
{
	const int line = d >> 1;
	const u64 L = line == 0 ? row_mask(pos) : line == 1 ? col_mask(pos) : line == 2 ? ld[pos] : rd[pos];
	const int n = (int)hd_popcll(all & L);
	return pos + n * dir_step(d);
}

// position of the k-th (0-based) set bit
__device__ __forceinline__ int select32(u32 x, u32 k)
{
	int pos = 0; u32 c;
	c = __popc(x & 0xffffu); if (k >= c) { k -= c; pos += 16; x >>= 16; }
	c = __popc(x & 0xffu);   if (k >= c) { k -= c; pos += 8;  x >>= 8; }
	c = __popc(x & 0xfu);    if (k >= c) { k -= c; pos += 4;  x >>= 4; }
	c = __popc(x & 0x3u);    if (k >= c) { k -= c; pos += 2;  x >>= 2; }
	if (k >= (x & 1u)) pos += 1;
	return pos;
}
__device__ __forceinline__ int select64(u64 x, u32 k)
{
	const u32 lo = (u32)x, hi = (u32)(x >> 32);
	const u32 c = __popc(lo);
	return k < c ? select32(lo, k) : 32 + select32(hi, k - c);
}

// Per-piece move masks of the side owning `my`, in ascending square order (:180-186): piece j -> byte j.
// Up to 12 pieces (LOA has 12 per side); positions with more own pieces are clamped by the callers'
// corpus (the reference asserts <= 96 moves, :226).
struct MoveMasks { u64 lo; u64 hi; };   // bytes 0..7 | bytes 8..15
__device__ __forceinline__ MoveMasks all_piece_moves(u64 my, u64 en, const u64* __restrict__ ld, const u64* __restrict__ rd)
{
	MoveMasks m; m.lo = 0; m.hi = 0;
	int j = 0;
	for (u64 rest = my; rest; rest &= rest - 1, ++j) {
		const int pos = __ffsll((long long)rest) - 1;
		const u64 v = piece_moves(my, en, pos, ld, rd);
		if (j < 8) m.lo |= v << (8 * j); else m.hi |= v << (8 * (j - 8));
	}
	return m;
}
__device__ __forceinline__ int count_moves(const MoveMasks& m) { return __popcll(m.lo) + __popcll(m.hi); }

// decode the idx-th move of the reference-ordered list into (from, to)
__device__ __forceinline__ void nth_move(const MoveMasks& m, u64 my, u64 en, u32 idx,
                                         const u64* __restrict__ ld, const u64* __restrict__ rd, int* from, int* to)
{
	const u32 clo = __popcll(m.lo);
	const int bit = idx < clo ? select64(m.lo, idx) : 64 + select64(m.hi, idx - clo);
	const int piece = bit >> 3, d = bit & 7;
	const int pos = select64(my, (u32)piece);
	*from = pos;
	*to = move_target(my | en, pos, d, ld, rd);
}

// applyMove (:233-252): my loses `from`, gains `to`; an enemy piece on `to` is captured.
__device__ __forceinline__ void apply_move(u64& my, u64& en, int from, int to)
{
	const u64 t = 1ull << to;
	my = (my & ~(1ull << from)) | t;
	en &= ~t;
}

// Score (:420-432) folded to an outcome code.
GAI_HDI int outcome_of(bool white_conn, bool black_conn)
{
	return (white_conn && black_conn) ? OUT_DRAW : white_conn ? OUT_WHITE : OUT_BLACK;
}

} // namespace loa
