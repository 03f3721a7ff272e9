// include/gameai/philox.h -- Philox4x32-10 choice stream shared by host C++ and device code.
//
// The reference has no counter-based generator: MC::Player draws from std::default_random_engine
// (MCTSPlayer.cpp:472-476), whose stream is implementation-defined.  The GPU rollout needs a stream
// that is a pure function of (seed, playout id, ply) so that a playout's result does not depend on
// which lane, warp or GPU ran it.  Stream definition:
//     key = (seed_lo, seed_hi); counter = (id_lo, id_hi, ply >> 2, 0); word = ply & 3
//     idx = (uint64(word) * n) >> 32
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GAI_HD __host__ __device__ __forceinline__
#else
#define GAI_HD inline
#endif

namespace gai {

struct Philox4 { uint32_t v[4]; };

GAI_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

GAI_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
	const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
	for (int r = 0; r < 10; ++r) {
		const uint32_t h0 = mulhi32(M0, c0), l0 = M0 * c0;
		const uint32_t h1 = mulhi32(M1, c2), l1 = M1 * c2;
		c0 = h1 ^ c1 ^ k0; c1 = l1;
		c2 = h0 ^ c3 ^ k1; c3 = l0;
		k0 += W0; k1 += W1;
	}
	Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
	return o;
}

// the four choice words for plies 4q .. 4q+3 of playout `id`
GAI_HD Philox4 choice_block(uint64_t key, uint64_t id, uint32_t q)
{
	return philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), q, 0u, (uint32_t)key, (uint32_t)(key >> 32));
}
GAI_HD uint32_t choice_index(uint32_t word, uint32_t n) { return mulhi32(word, n); }

} // namespace gai
