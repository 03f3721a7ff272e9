// profiles/ubench/pipes.cu -- which issue pipe does an integer instruction use on sm_100a, and at what rate?
//
// The rollout kernel is bound by the ALU pipe (LOP3/SHF/PRMT/IADD3: one warp instruction every 2 cycles per SM
// sub-partition) while the FMA pipe idles.  This microbenchmark measures, per candidate instruction, the sustained
// warp-instructions per cycle per sub-partition (a) alone and (b) interleaved 1:1 with LOP3: if the mix runs at the
// SUM of the two rates the instruction does not share the ALU pipe and work can be moved onto it for free.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes > pipes.json
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16

enum Op { LOP3, PRMT, SHF, IADD3, IMAD, IMADHI, IMADWIDE, DP4A, DP2A, POPC, SEL, SHL, FFMA, LDS_, BREV, FLO, IMNMX, ISETP, VABSDIFF, LEA, NOPS };
static const char* op_name[] = { "lop3", "prmt", "shf", "lop3+iadd3", "imad", "imad.hi", "imad.wide", "dp4a", "dp2a", "popc", "sel", "shl", "ffma", "lds", "brev", "flo", "imnmx", "isetp+sel", "vabsdiff4", "lea(a*8+b)" };

template <int OP> __device__ __forceinline__ uint32_t one(uint32_t x, uint32_t a, uint32_t b, const uint32_t* sm)
{
	uint32_t r;
	if (OP == LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == PRMT) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == IADD3) { asm volatile("{.reg .u32 t; xor.b32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(r) : "r"(x), "r"(a), "r"(b)); }   /* LOP3 + IADD3 */
	else if (OP == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == IMADHI) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == IMADWIDE) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(a)); r = (uint32_t)(w >> 32); }
	else if (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == DP2A) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == POPC) asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(x));
	else if (OP == SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %3, 0; selp.u32 %0, %1, %2, p;}" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == SHL) asm volatile("shl.b32 %0, %1, 3;" : "=r"(r) : "r"(x));
	else if (OP == FFMA) { float f; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(__uint_as_float(x)), "f"(__uint_as_float(a)), "f"(__uint_as_float(b))); r = __float_as_uint(f); }
	else if (OP == LDS_) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"((x & 0x3fcu) + (uint32_t)__cvta_generic_to_shared(sm)));
	else if (OP == BREV) asm volatile("brev.b32 %0, %1;" : "=r"(r) : "r"(x));
	else if (OP == FLO) asm volatile("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
	else if (OP == IMNMX) asm volatile("max.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
	else if (OP == ISETP) asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; selp.u32 %0, %1, %3, p;}" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == VABSDIFF) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
	else if (OP == LEA) asm volatile("mad.lo.u32 %0, %1, 8, %2;" : "=r"(r) : "r"(x), "r"(a));
	else r = x;
	return r;
}

// A alone (B == NOPS) or A and B interleaved 1:1, CHAINS independent dependency chains per thread
template <int A, int B>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b, int iters)
{
	__shared__ uint32_t sm[256];
	sm[threadIdx.x] = threadIdx.x * 4u;
	__syncthreads();
	uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x * 4u + c; y[c] = threadIdx.x + 77u * c; }
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
			for (int c = 0; c < CHAINS; ++c) {
				x[c] = one<A>(x[c], a, b, sm);
				if (B != NOPS) y[c] = one<B>(y[c], a, b, sm);
			}
		}
	}
	uint32_t s = 0;
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) s ^= x[c] ^ y[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int A, int B> double run(uint32_t* d_out, int sms, double mhz)
{
	const int iters = 2000, blocks = sms * 4, threads = 256;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<A, B><<<blocks, threads>>>(d_out, 0x01020304u, 0x5410u, 10);
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	k<A, B><<<blocks, threads>>>(d_out, 0x01020304u, 0x5410u, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	const double per_thread = (double)iters * UNROLL * CHAINS * (B == NOPS ? 1 : 2);
	const double warp_inst = per_thread * blocks * threads / 32.0;
	return warp_inst / (ms * 1e-3) / (sms * 4.0 * mhz * 1e6);       // warp-inst per cycle per sub-partition
}

template <int A> void both(uint32_t* d_out, int sms, double mhz, bool last)
{
	const double alone = run<A, NOPS>(d_out, sms, mhz);
	const double mix = run<A, LOP3>(d_out, sms, mhz);
	const double mixf = run<A, IMAD>(d_out, sms, mhz);
	printf("  {\"op\": \"%s\", \"alone\": %.3f, \"with_lop3\": %.3f, \"with_imad\": %.3f}%s\n", op_name[A], alone, mix, mixf, last ? "" : ",");
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double mhz = khz / 1000.0;
	uint32_t* d_out; cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 4 * 256 * 4);
	printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_nominal\": %.0f, \"unit\": \"warp-instructions / cycle / sub-partition at the nominal clock\", \"ops\": [\n", p.name, p.multiProcessorCount, mhz);
	both<LOP3>(d_out, p.multiProcessorCount, mhz, false);
	both<PRMT>(d_out, p.multiProcessorCount, mhz, false);
	both<SHF>(d_out, p.multiProcessorCount, mhz, false);
	both<IADD3>(d_out, p.multiProcessorCount, mhz, false);
	both<IMAD>(d_out, p.multiProcessorCount, mhz, false);
	both<IMADHI>(d_out, p.multiProcessorCount, mhz, false);
	both<IMADWIDE>(d_out, p.multiProcessorCount, mhz, false);
	both<DP4A>(d_out, p.multiProcessorCount, mhz, false);
	both<DP2A>(d_out, p.multiProcessorCount, mhz, false);
	both<POPC>(d_out, p.multiProcessorCount, mhz, false);
	both<SEL>(d_out, p.multiProcessorCount, mhz, false);
	both<SHL>(d_out, p.multiProcessorCount, mhz, false);
	both<FFMA>(d_out, p.multiProcessorCount, mhz, false);
	both<LDS_>(d_out, p.multiProcessorCount, mhz, false);
	both<BREV>(d_out, p.multiProcessorCount, mhz, false);
	both<FLO>(d_out, p.multiProcessorCount, mhz, false);
	both<IMNMX>(d_out, p.multiProcessorCount, mhz, false);
	both<ISETP>(d_out, p.multiProcessorCount, mhz, false);
	both<VABSDIFF>(d_out, p.multiProcessorCount, mhz, false);
	both<LEA>(d_out, p.multiProcessorCount, mhz, true);
	printf("]}\n");
	return 0;
}
