// game-ai_b200/csrc/loa_fast.cuh -- LOA rules, "rotated boards + line LUT" formulation (device).
//
// Same observable behaviour as loa_rules.cuh / LinesOfActionZasady.cpp:136-232 (move ORDER included), but
// organised for the B200 SM: the path is ALU-pipe bound (LOP3/SHF/IADD issue at half rate), so per piece the
// work is four shared-memory lookups instead of ~150 bit operations:
//
//  * each colour is held in four 64-bit "rotated" boards in registers, so that every line through a square is
//    ONE BYTE of one register pair:
//        R  : byte = row r          cell = column c           (the plain bitboard, LinesOfActionZasady.cpp:43)
//        C  : byte = column c       cell = row r              (transposed)
//        D1 : byte = (r - c) & 7    cell = row r              ("left" diagonal A1->H8, :93-113)
//        D2 : byte = (r + c) & 7    cell = row r              ("right" diagonal H1->A8, :114-135)
//    In R, D1 and D2 the reference's FIRST direction of the line (W, SW, SE -- :194-223) is "towards lower cells" and
//    the second (E, NE, NW) "towards higher cells"; in C it is the other way round (N = higher row comes first), the
//    two bits of a C field are swapped when a piece's move mask is assembled.  C, D1 and D2 all index their cells by
//    ROW, which is what the static sweep in loa_sweep.cuh relies on (its per-row totals are a vertical sum).
//    A D1/D2 byte holds two complementary diagonals (lengths k and 8-k); the cells of the other diagonal are
//    presented to the LUT as "phantom" cells.
//  * one table LUT[(o8, e8)] (own / enemy occupancy of the 8 cells) -> 16 bits, 2 per cell:
//    bit 2p = own piece on cell p may move towards lower cells, bit 2p+1 = towards higher cells, by the
//    number of pieces on the line (:191-192), not landing on an own piece (:142), not jumping an enemy
//    (:146-148).  A cell with BOTH o and e set is a phantom: not counted, not a mover, cannot be landed on.
//    Index = o8 + 259*e8.  The stride decides the shared-memory bank of an entry, ((o + 259 e) >> 1) & 31: an ODD stride
//    lets bit 0 of e reach the bank bits through the carry; simulated on reachable positions (32 random lanes, the 30
//    lookups of lines longer than 4 cells) 259 costs 3.7 wavefronts per lookup, the former 258 4.8, 256 9.7.
//  * lines of at most 4 cells (the short diagonals: 12 of the 42 lookups) live in ONE nibble of their byte; the big table
//    hashes them badly (their varying bits fall outside the bank bits: 7-8 wavefronts).  They use a 256-entry table
//    indexed by the two nibbles, replicated once per lane so that lane L only ever touches bank L: no conflicts at all.
#pragma once
#include "loa_rules.cuh"
#include "loa_euler.cuh"

namespace loa {

#ifndef GAI_LUT_STRIDE
#define GAI_LUT_STRIDE 259
#endif
constexpr int LUT_STRIDE = GAI_LUT_STRIDE;        // >= 256 and odd (see above); 2 * stride = 14 * 37 keeps the enemy weight a dp4a byte
constexpr int LUT_ENTRIES = LUT_STRIDE * 255 + 256;          // 66046
constexpr int LUT_BYTES = ((LUT_ENTRIES * 2 + 15) / 16) * 16; // 132608
// two 256-entry u32 tables follow the line LUT in the global image (expanded to one copy per lane in shared memory)
constexpr int SLT_OFF = LUT_BYTES, TPC_OFF = LUT_BYTES + 1024, LUT_IMAGE_BYTES = LUT_BYTES + 2048;

// One LUT entry, by definition (host + device; the table is built once per context on the host).
__host__ __device__ inline uint16_t lut_entry(u32 o, u32 e)
{
	const u32 own = o & ~e, enemy = e & ~o;
	int n = 0;
	for (int i = 0; i < 8; ++i) n += ((own | enemy) >> i) & 1u;
	u32 res = 0;
	for (int p = 0; p < 8; ++p) {
		if (!((own >> p) & 1u)) continue;
		const int tu = p + n, td = p - n;
		if (tu < 8 && !((o >> tu) & 1u)) {                 // on the line, not own, not phantom
			bool blocked = false;
			for (int q = p + 1; q < tu; ++q) blocked |= (enemy >> q) & 1u;
			if (!blocked) res |= 2u << (2 * p);
		}
		if (td >= 0 && !((o >> td) & 1u)) {
			bool blocked = false;
			for (int q = td + 1; q < p; ++q) blocked |= (enemy >> q) & 1u;
			if (!blocked) res |= 1u << (2 * p);
		}
	}
	return (uint16_t)res;
}

// short-line table: 4 cells, index = own nibble + 16 * enemy nibble (a cell set in both = phantom) -> 8 bits, 2 per cell
__host__ __device__ inline u32 slt_entry(u32 idx) { return (u32)lut_entry((idx & 15u) | 0xf0u, (idx >> 4) | 0xf0u) & 0xffu; }
// field-count table: a byte of four 2-bit fields -> the popcount of each field in its own byte lane
__host__ __device__ inline u32 tpc_entry(u32 b)
{
	u32 r = 0;
	for (int j = 0; j < 4; ++j) { const u32 f = (b >> (2 * j)) & 3u; r |= ((f & 1u) + (f >> 1)) << (8 * j); }
	return r;
}

// rotated-board bit positions of square s = 8r + c
__host__ __device__ constexpr int pos_c(int s)  { return 8 * (s & 7) + (s >> 3); }
__host__ __device__ constexpr int pos_d1(int s) { return 8 * (((s >> 3) - (s & 7)) & 7) + (s >> 3); }
__host__ __device__ constexpr int pos_d2(int s) { return 8 * (((s >> 3) + (s & 7)) & 7) + (s >> 3); }
// phantom cells of the D1 / D2 byte seen from square s (cells that belong to the complementary diagonal)
__host__ __device__ constexpr u32 phantom_d1(int s)
{
	const int r = s >> 3, c = s & 7, k = (r - c) & 7;           // r >= c: the diagonal r - c = k holds rows k..7 of the byte
	return (r >= c) ? ((1u << k) - 1u) : ((0xffu << k) & 0xffu);   // r <  c: the diagonal r - c = k - 8 holds rows 0..k-1
}
__host__ __device__ constexpr u32 phantom_d2(int s)
{
	const int r = s >> 3, c = s & 7, k = (r + c) & 7;
	return (r + c <= 7) ? ((0xffu << (k + 1)) & 0xffu) : ((1u << (k + 1)) - 1u);
}

// per-square constants, one 64-bit word:  lo = phD1 | phD2<<8 | byteD1<<16 | byteD2<<24 ; hi = posC | posD1<<8 | posD2<<16
struct SquareTable { u64 w[64]; };
__host__ __device__ constexpr SquareTable make_square_table()
{
	SquareTable t{};
	for (int s = 0; s < 64; ++s) {
		const u64 lo = (u64)phantom_d1(s) | ((u64)phantom_d2(s) << 8) | ((u64)(pos_d1(s) >> 3) << 16) | ((u64)(pos_d2(s) >> 3) << 24);
		const u64 hi = (u64)pos_c(s) | ((u64)pos_d1(s) << 8) | ((u64)pos_d2(s) << 16);
		t.w[s] = lo | (hi << 32);
	}
	return t;
}

// line masks by (line type, square): [0] row, [1] column, [2] left diagonal, [3] right diagonal (LinesOfActionZasady.cpp:82-135),
// and the 8-neighbourhood of every square (:59-81); 2.5 KB, staged into shared memory by the sweep kernels
struct LineTable { u64 line[4][64]; u64 nbr[64]; };
__host__ __device__ constexpr LineTable make_line_table()
{
	LineTable t{};
	for (int s = 0; s < 64; ++s) {
		t.line[0][s] = row_mask(s); t.line[1][s] = col_mask(s); t.line[2][s] = ldiag_mask_slow(s); t.line[3][s] = rdiag_mask_slow(s);
		const int r = s >> 3, c = s & 7;
		u64 m = 0;
		for (int dr = -1; dr <= 1; ++dr) for (int dc = -1; dc <= 1; ++dc) {
			const int rr = r + dr, cc = c + dc;
			if ((dr || dc) && rr >= 0 && rr < 8 && cc >= 0 && cc < 8) m |= 1ull << (rr * 8 + cc);
		}
		t.nbr[s] = m;
	}
	return t;
}

struct Side { u64 R, C, D1, D2; };      // one colour, four rotations

GAI_HDI u64 bit64(int p) { return 1ull << p; }

// build the rotations of one colour from its plain bitboard (once per playout)
GAI_HDI Side make_side(u64 b, const u64* __restrict__ sqt)
{
	Side s; s.R = b; s.C = 0; s.D1 = 0; s.D2 = 0;
	for (u64 rest = b; rest; rest &= rest - 1) {
		const int q = hd_ffsll(rest) - 1;
		const u32 hi = (u32)(sqt[q] >> 32);
		s.C |= bit64(hi & 63u); s.D1 |= bit64((hi >> 8) & 63u); s.D2 |= bit64((hi >> 16) & 63u);
	}
	return s;
}

__device__ __forceinline__ u32 byte_of(u64 x, u32 i) { return __byte_perm((u32)x, (u32)(x >> 32), i); }   // byte i in bits 0..7, junk above

// LUT address (in bytes) of the line with own/enemy occupancy bytes po/pe (junk above bit 7) and phantom cells ph
__device__ __forceinline__ u32 lut_lookup(const uint16_t* __restrict__ lut, u32 po, u32 pe, u32 ph)
{
	const u32 o = (po & 0xffu) | ph, e = (pe & 0xffu) | ph;
	return lut[o + e * (u32)LUT_STRIDE];
}

// 8-bit move mask (bit d = reference direction d legal) of the own piece on square s
__device__ __forceinline__ u32 piece_moves_fast(const Side& my, const Side& en, int s, const uint16_t* __restrict__ lut,
                                                const u64* __restrict__ sqt)
{
	const u32 r = (u32)s >> 3, c = (u32)s & 7u;
	const u32 t = (u32)sqt[s];
	const u32 ph1 = __byte_perm(t, 0, 0x4440), ph2 = __byte_perm(t, 0, 0x4441);
	const u32 k1 = __byte_perm(t, 0, 0x4442), k2 = __byte_perm(t, 0, 0x4443);
	const u32 vR = lut_lookup(lut, byte_of(my.R, r), byte_of(en.R, r), 0u);
	const u32 vC = lut_lookup(lut, byte_of(my.C, c), byte_of(en.C, c), 0u);
	const u32 v1 = lut_lookup(lut, byte_of(my.D1, k1), byte_of(en.D1, k1), ph1);
	const u32 v2 = lut_lookup(lut, byte_of(my.D2, k2), byte_of(en.D2, k2), ph2);
	const u32 xR = (vR >> (2u * c)) & 3u;            // W, E
	const u32 xCs = (vC >> (2u * r)) & 3u;           // bit0 = S (lower row), bit1 = N
	const u32 xC = ((xCs >> 1) | (xCs << 1)) & 3u;   // N, S
	const u32 x1 = (v1 >> (2u * r)) & 3u;            // SW, NE
	const u32 x2 = (v2 >> (2u * r)) & 3u;            // SE, NW
	return xR + xC * 4u + x1 * 16u + x2 * 64u;
}

__device__ __forceinline__ MoveMasks all_piece_moves_fast(const Side& my, const Side& en, const uint16_t* __restrict__ lut,
                                                          const u64* __restrict__ sqt)
{
	MoveMasks m; m.lo = 0; m.hi = 0;
	int j = 0;
	for (u64 rest = my.R; rest; rest &= rest - 1, ++j) {
		const int s = __ffsll((long long)rest) - 1;
		const u64 v = piece_moves_fast(my, en, s, lut, sqt);
		if (j < 8) m.lo |= v << (8 * j); else m.hi |= v << (8 * (j - 8));
	}
	return m;
}

// move the own piece from -> to in all four rotations; captures the enemy piece on `to` if any
GAI_HDI void apply_move_fast(Side& my, Side& en, int from, int to, const u64* __restrict__ sqt)
{
	const u32 hf = (u32)(sqt[from] >> 32), ht = (u32)(sqt[to] >> 32);
	const u64 tR = bit64(to), tC = bit64(ht & 63u), t1 = bit64((ht >> 8) & 63u), t2 = bit64((ht >> 16) & 63u);
	my.R = (my.R ^ bit64(from)) | tR;
	my.C = (my.C ^ bit64(hf & 63u)) | tC;
	my.D1 = (my.D1 ^ bit64((hf >> 8) & 63u)) | t1;
	my.D2 = (my.D2 ^ bit64((hf >> 16) & 63u)) | t2;
	en.R &= ~tR; en.C &= ~tC; en.D1 &= ~t1; en.D2 &= ~t2;
}

// target square through the line table: one 8-byte load instead of a branch per line type
GAI_HDI int move_target_tab(u64 all, int pos, int d, const u64* __restrict__ lines)
{
	const u64 L = lines[(d >> 1) * 64 + pos];
	return pos + (int)hd_popcll(all & L) * dir_step(d);
}
// apply_move_fast with the three rotated positions unpacked by IDP.4A (one-byte weight = byte extract on the FMA pipe);
// returns true when the captured piece (if any) had no neighbour of its own colour -- the only case in which losing it
// can leave that colour connected: it was a group of its own, so before the capture there were >= 2 groups and every
// other capture leaves >= 2 (the rest of the victim's group + the untouched ones).
GAI_HDI bool apply_move_sweep(Side& my, Side& en, int from, int to, const u64* __restrict__ sqt, const u64* __restrict__ nbr)
{
	const u32 hf = (u32)(sqt[from] >> 32), ht = (u32)(sqt[to] >> 32);
	const u64 tR = bit64(to), tC = bit64((int)hd_dp4a(ht, 0x00000001u, 0u)), t1 = bit64((int)hd_dp4a(ht, 0x00000100u, 0u)), t2 = bit64((int)hd_dp4a(ht, 0x00010000u, 0u));
	const bool cap = (en.R & tR) != 0;
	my.R = (my.R ^ bit64(from)) | tR;
	my.C = (my.C ^ bit64((int)hd_dp4a(hf, 0x00000001u, 0u))) | tC;
	my.D1 = (my.D1 ^ bit64((int)hd_dp4a(hf, 0x00000100u, 0u))) | t1;
	my.D2 = (my.D2 ^ bit64((int)hd_dp4a(hf, 0x00010000u, 0u))) | t2;
	en.R &= ~tR; en.C &= ~tC; en.D1 &= ~t1; en.D2 &= ~t2;
	return cap && (en.R & nbr[to]) == 0;
}

// apply for the sweep kernel with the Euler characteristics of both colours kept up to date (loa_euler.cuh): the mover's
// figure loses `from` and gains `to`, the other colour loses `to` on a capture.  Returns true on a capture.
// hf / ht: the rotated positions of from / to (high word of the square table); wf / wt: their window controls (loa_euler.cuh).
// The caller loads them: the rollout kernel from per-lane copies of the two tables (conflict-free), everything else from
// the plain 64-entry tables through the overload below.
GAI_HDI bool apply_move_chi(Side& my, Side& en, int& chi_my, int& chi_en, int from, int to, u32 hf, u32 ht, u64 wf, u64 wt,
                            const signed char* __restrict__ chitab)
{
	const u64 tR = bit64(to), tC = bit64((int)hd_dp4a(ht, 0x00000001u, 0u)), t1 = bit64((int)hd_dp4a(ht, 0x00000100u, 0u)), t2 = bit64((int)hd_dp4a(ht, 0x00010000u, 0u));
	const bool cap = (en.R & tR) != 0;
	const int d_from = chitab[chi_index(my.R, wf)];               // the centre is masked out of its own neighbourhood
	const u64 r1 = my.R ^ bit64(from);
	const int d_to = chitab[chi_index(r1, wt)];
	const int d_cap = chitab[chi_index(en.R, wt)];
	chi_my += d_to - d_from;
	chi_en -= cap ? d_cap : 0;
	my.R = r1 | tR;
	my.C = (my.C ^ bit64((int)hd_dp4a(hf, 0x00000001u, 0u))) | tC;
	my.D1 = (my.D1 ^ bit64((int)hd_dp4a(hf, 0x00000100u, 0u))) | t1;
	my.D2 = (my.D2 ^ bit64((int)hd_dp4a(hf, 0x00010000u, 0u))) | t2;
	en.R &= ~tR; en.C &= ~tC; en.D1 &= ~t1; en.D2 &= ~t2;
	return cap;
}
GAI_HDI bool apply_move_chi(Side& my, Side& en, int& chi_my, int& chi_en, int from, int to, const u64* __restrict__ sqt,
                            const u64* __restrict__ nbx, const signed char* __restrict__ chitab)
{
	return apply_move_chi(my, en, chi_my, chi_en, from, to, (u32)(sqt[from] >> 32), (u32)(sqt[to] >> 32), nbx[from], nbx[to], chitab);
}

} // namespace loa
