// game-ai_b200/csrc/gameai_gpu.cu -- C ABI of libgameai_gpu.so (see include/gameai_gpu.h).
// Context = one device, one stream, pinned staging + device scratch grown on demand.  No CPU fallback.
#include "gameai_gpu.h"
#include <algorithm>
#include "kernels.h"
#include "gameai/philox.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <mutex>

namespace {

thread_local char g_create_error[512] = "";

struct Buf {            // grow-only device or pinned-host buffer
	void* p = nullptr; size_t cap = 0; bool host = false;
};

// minimal NCCL surface, resolved at run time (dlopen) so the library loads without NCCL installed
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
	void* h = nullptr;
	int (*GetUniqueId)(void*) = nullptr;
	int (*CommInitRank)(ncclComm_t*, int, /* ncclUniqueId by value: 128 bytes */ struct Id128, int) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
};
struct Id128 { char b[128]; };

} // namespace

// one queued batch of the ticketed API: own staging + device buffers, so that batch k+1 can be uploaded and queued
// while batch k plays out
struct Slot {
	Buf d_in, d_off, d_boards, d_side, d_stm, d_out, d_prep, d_misc;
	Buf h_in, h_off, h_out, h_stat;
	cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
	bool busy = false, direct = false, children = false;
	gai_loa_result* out = nullptr; uint32_t n_out = 0; unsigned gen = 0;
};

constexpr int GAI_PAN_MAX_CHUNKS = 32;
constexpr uint32_t GAI_PAN_CHUNK = 65536;          // deals per pipeline stage (2.6 MB in, 3 MB out, ~0.4 ms of kernel at depth 50)

struct gai_ctx {
	int device = 0;
	Slot slots[GAI_MAX_INFLIGHT];
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
	cudaDeviceProp prop;
	char err[512];
	uint64_t launches = 0;
	int threads = 256, blocks_per_sm = 4;   // baseline kernel launch shape
	int impl = 2;                           // 0 = bitboard rules (loa_rules.cuh), 1 = rotated boards + line LUT per piece (loa_fast.cuh), 2 = static line sweep (loa_sweep.cuh)
	int fast_threads = 1024;                // one CTA per SM for the LUT kernels
	int sweep_threads = 896;                // the static-sweep kernel fits 72 registers without spilling: 28 warps per SM
	int sweep_unroll = 2;                   // plies per loop trip of the static-sweep kernel (1, 2, 4); GAI_SWEEP_UNROLL overrides
	void* d_lut = nullptr;
	// scratch
	Buf d_in, d_in2, d_out, d_out2, d_out3, d_boards, d_stm, d_misc, d_prep;
	Buf h_in, h_in2, h_out, h_out2, h_out3;
	// async rollout bookkeeping
	gai_loa_result* pending_out = nullptr; uint32_t pending_n = 0; bool pending_direct = false;   // direct: results were DMA-ed into the caller's pinned buffer
	// Pan host-buffer pipeline (gai_pan_rollout): copy-in and copy-out streams beside the compute stream, events per chunk
	cudaStream_t pan_in = nullptr, pan_out = nullptr;
	cudaEvent_t pan_ev[4][GAI_PAN_MAX_CHUNKS] = {};      // [0] chunk uploaded, [1] kernels start, [2] kernels end, [3] chunk downloaded
	Buf pan_misc;                                         // 64 bytes per chunk: {playouts, plies, work counter}, then the bad-state flag
	// nccl
	NcclApi nccl; ncclComm_t comm = nullptr; int rank = 0, world = 1;
};

extern "C" int gai_sync(gai_ctx* ctx);

namespace {

int fail(gai_ctx* c, int code, const char* fmt, const char* a = "", const char* b = "")
{
	if (c) snprintf(c->err, sizeof c->err, fmt, a, b);
	else snprintf(g_create_error, sizeof g_create_error, fmt, a, b);
	return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, GAI_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

int ensure(gai_ctx* ctx, Buf& b, size_t bytes, bool host)
{
	if (bytes <= b.cap) return GAI_OK;
	size_t cap = bytes + bytes / 4 + 256;
	if (b.p) { if (b.host) cudaFreeHost(b.p); else cudaFree(b.p); b.p = nullptr; b.cap = 0; }
	b.host = host;
	cudaError_t e = host ? cudaMallocHost(&b.p, cap) : cudaMalloc(&b.p, cap);
	if (e != cudaSuccess) { b.p = nullptr; return fail(ctx, GAI_ERR_NOMEM, "%s of scratch failed: %s", host ? "cudaMallocHost" : "cudaMalloc", cudaGetErrorString(e)); }
	b.cap = cap;
	return GAI_OK;
}
#define EN(buf, bytes, host) do { int r_ = ensure(ctx, buf, (bytes), host); if (r_) return r_; } while (0)

void release(Buf& b) { if (b.p) { if (b.host) cudaFreeHost(b.p); else cudaFree(b.p); } b.p = nullptr; b.cap = 0; }

int check_launch(gai_ctx* ctx, const char* what)
{
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(ctx, GAI_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
	ctx->launches += 1;
	return GAI_OK;
}
#define LAUNCHED(what) do { int r_ = check_launch(ctx, what); if (r_) return r_; } while (0)

int rollout_grid(const gai_ctx* ctx, uint64_t total_items)
{
	int blocks = ctx->prop.multiProcessorCount * ctx->blocks_per_sm;      // persistent: a multiple of the SM count
	const uint64_t need = (total_items + ctx->threads - 1) / ctx->threads;
	if (need < (uint64_t)blocks) blocks = (int)(need ? need : 1);
	return blocks;
}

// enqueue: zero out/stat/counter, rollout kernel.  d_misc layout: [0..1] stats u64, [2] work counter.
int enqueue_rollout(gai_ctx* ctx, Buf& misc, Buf& prep, cudaEvent_t ev_k0, cudaEvent_t ev_k1, const void* d_boards, const uint32_t* d_stm, uint32_t n_leaves, uint32_t ppl,
                    uint32_t max_plies, uint64_t key, uint64_t first_id, void* d_out)
{
	EN(misc, 128, false);
	// d_prep: [n x 64 B rotations][n x u32 order][n x u8 piece counts][64 x u32 histogram + cursors]
	const size_t prep_b = (size_t)n_leaves * 64, ord_b = (size_t)n_leaves * 4, pc_b = ((size_t)n_leaves + 255) / 256 * 256;
	const size_t chi_b = (size_t)n_leaves * 2;
	if (n_leaves >= (1u << 30)) return fail(ctx, GAI_ERR_INVALID, "at most 2^30 - 1 leaves per batch");
	if (ctx->impl >= 1) EN(prep, prep_b + ord_b + pc_b + 256 + chi_b, false);
	CU(cudaMemsetAsync(misc.p, 0, 64, ctx->stream));
	CU(cudaMemsetAsync(d_out, 0, (size_t)n_leaves * sizeof(gai_loa_result), ctx->stream));
	const uint64_t total = (uint64_t)n_leaves * ppl;
	CU(cudaEventRecord(ev_k0, ctx->stream));
	if (ctx->impl >= 1) {
		// rotations of every leaf, once per batch (inside the timed region: it is part of the step)
		char* pb = (char*)prep.p;
		uint32_t* d_order = (uint32_t*)(pb + prep_b); uint8_t* d_pc = (uint8_t*)(pb + prep_b + ord_b); uint32_t* d_hist = (uint32_t*)(pb + prep_b + ord_b + pc_b);
		uint16_t* d_chi = (uint16_t*)(pb + prep_b + ord_b + pc_b + 256);
		CU(cudaMemsetAsync(d_hist, 0, 256, ctx->stream));
		launch_loa_prepare_leaves(d_boards, n_leaves, pb, d_pc, d_hist, d_order, d_chi, ctx->stream);
		LAUNCHED("k_prepare_leaves"); ctx->launches += 1;      // + k_order_leaves
		launch_loa_rollout_fast(ctx->d_lut, pb, d_stm, d_order, d_chi, n_leaves, ppl, max_plies, key, first_id, (uint32_t*)d_out,
		                        (uint64_t*)misc.p, (unsigned long long*)misc.p + 2,
		                        ctx->impl == 2 ? ctx->sweep_threads : ctx->fast_threads, ctx->prop.multiProcessorCount, ctx->impl == 2 ? ctx->sweep_unroll : 0, ctx->stream);
	} else
		launch_loa_rollout(d_boards, d_stm, n_leaves, ppl, max_plies, key, first_id, (uint32_t*)d_out,
		                   (uint64_t*)misc.p, (unsigned long long*)misc.p + 2,
		                   ctx->threads, rollout_grid(ctx, total), ctx->stream);
	LAUNCHED("k_rollout");
	CU(cudaEventRecord(ev_k1, ctx->stream));
	return GAI_OK;
}
int enqueue_rollout(gai_ctx* ctx, const void* d_boards, const uint32_t* d_stm, uint32_t n_leaves, uint32_t ppl,
                    uint32_t max_plies, uint64_t key, uint64_t first_id, void* d_out)
{
	return enqueue_rollout(ctx, ctx->d_misc, ctx->d_prep, ctx->ev[1], ctx->ev[2], d_boards, d_stm, n_leaves, ppl, max_plies, key, first_id, d_out);
}

int tickets_out(const gai_ctx* ctx) { int k = 0; for (const Slot& sl : ctx->slots) k += sl.busy ? 1 : 0; return k; }
// The synchronous entry points share the context's scratch buffers with each other and with gai_loa_rollout_async:
// finish a pending async batch first, refuse to run while ticketed batches are queued.
int quiesce(gai_ctx* ctx, const char* who)
{
	if (tickets_out(ctx)) return fail(ctx, GAI_ERR_INVALID, "%s: ticketed batches are outstanding on this context (gai_wait them first)", who);
	if (ctx->pending_out) return gai_sync(ctx);
	return GAI_OK;
}
#define QUIESCE(who) do { int q_ = quiesce(ctx, who); if (q_) return q_; } while (0)

int fill_stats(gai_ctx* ctx, gai_rollout_stats* stats, bool have_total)
{
	if (!stats) return GAI_OK;
	uint64_t h[2];
	CU(cudaMemcpy(h, ctx->d_misc.p, sizeof h, cudaMemcpyDeviceToHost));
	stats->playouts = h[0]; stats->plies = h[1];
	CU(cudaEventElapsedTime(&stats->kernel_ms, ctx->ev[1], ctx->ev[2]));
	stats->total_ms = stats->kernel_ms;
	if (have_total) CU(cudaEventElapsedTime(&stats->total_ms, ctx->ev[0], ctx->ev[3]));
	return GAI_OK;
}

} // namespace

extern "C" {

int gai_abi_version(void) { return GAI_ABI_VERSION; }

void gai_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	const gai::Philox4 r = gai::philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
	for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

int gai_create(int device, gai_ctx** out)
{
	if (!out) return fail(nullptr, GAI_ERR_INVALID, "gai_create: out is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(nullptr, GAI_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	if (device < 0 || device >= count) return fail(nullptr, GAI_ERR_INVALID, "gai_create: device ordinal out of range");
	gai_ctx* ctx = new gai_ctx();
	if (const char* u = getenv("GAI_SWEEP_UNROLL")) { const int v = atoi(u); if (v == 1 || v == 2 || v == 4) ctx->sweep_unroll = v; }      // tuning knob
	if (const char* u = getenv("GAI_SWEEP_THREADS")) { const int v = atoi(u); if (v >= 512 && v <= 1024 && v % 128 == 0) ctx->sweep_threads = v; }
	ctx->err[0] = 0;
	ctx->device = device;
	if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) {
		fail(nullptr, GAI_ERR_CUDA, "cudaSetDevice/GetDeviceProperties: %s", cudaGetErrorString(e)); delete ctx; return GAI_ERR_CUDA;
	}
	if (ctx->prop.major != 10) {
		fail(nullptr, GAI_ERR_NO_DEVICE, "device %s is not sm_100 (kernels are built for sm_100a only)", ctx->prop.name); delete ctx; return GAI_ERR_NO_DEVICE;
	}
	if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
		fail(nullptr, GAI_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); delete ctx; return GAI_ERR_CUDA;
	}
	for (auto& ev : ctx->ev) cudaEventCreate(&ev);
	for (Slot& sl : ctx->slots) for (auto& ev : sl.ev) cudaEventCreate(&ev);
	{	// line LUT of the fast rules: built on the host from its definition, resident in HBM/L2, staged to shared memory per CTA
		const int lb = loa_fast_lut_bytes();
		uint16_t* h = (uint16_t*)malloc(lb);
		loa_fast_build_lut_host(h);
		e = cudaMalloc(&ctx->d_lut, lb);
		if (e == cudaSuccess) e = cudaMemcpy(ctx->d_lut, h, lb, cudaMemcpyHostToDevice);
		free(h);
		if (e == cudaSuccess) e = loa_fast_prepare();
		if (e != cudaSuccess) { fail(nullptr, GAI_ERR_CUDA, "line-LUT setup: %s", cudaGetErrorString(e)); gai_destroy(ctx); return GAI_ERR_CUDA; }
	}
	*out = ctx;
	return GAI_OK;
}

void gai_destroy(gai_ctx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
	for (Buf* b : { &ctx->d_in, &ctx->d_in2, &ctx->d_out, &ctx->d_out2, &ctx->d_out3, &ctx->d_boards, &ctx->d_stm, &ctx->d_misc, &ctx->d_prep,
	                &ctx->h_in, &ctx->h_in2, &ctx->h_out, &ctx->h_out2, &ctx->h_out3 }) release(*b);
	for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
	for (Slot& sl : ctx->slots) {
		for (Buf* b : { &sl.d_in, &sl.d_off, &sl.d_boards, &sl.d_side, &sl.d_stm, &sl.d_out, &sl.d_prep, &sl.d_misc, &sl.h_in, &sl.h_off, &sl.h_out, &sl.h_stat }) release(*b);
		for (auto& ev : sl.ev) if (ev) cudaEventDestroy(ev);
	}
	if (ctx->d_lut) cudaFree(ctx->d_lut);
	release(ctx->pan_misc);
	for (auto& row : ctx->pan_ev) for (auto& ev : row) if (ev) cudaEventDestroy(ev);
	if (ctx->pan_in) cudaStreamDestroy(ctx->pan_in);
	if (ctx->pan_out) cudaStreamDestroy(ctx->pan_out);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

const char* gai_last_error(const gai_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int gai_get_device_info(gai_ctx* ctx, gai_device_info* info)
{
	if (!ctx || !info) return GAI_ERR_INVALID;
	memset(info, 0, sizeof *info);
	info->device = ctx->device;
	info->sm_count = ctx->prop.multiProcessorCount;
	info->cc_major = ctx->prop.major; info->cc_minor = ctx->prop.minor;
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
	info->clock_khz = khz;
	info->total_mem = ctx->prop.totalGlobalMem;
	strncpy(info->name, ctx->prop.name, sizeof info->name - 1);
	return GAI_OK;
}

int gai_sync(gai_ctx* ctx)
{
	if (!ctx) return GAI_ERR_INVALID;
	CU(cudaStreamSynchronize(ctx->stream));
	if (ctx->pending_out) {
		if (!ctx->pending_direct) memcpy(ctx->pending_out, ctx->h_out.p, (size_t)ctx->pending_n * sizeof(gai_loa_result));
		ctx->pending_out = nullptr; ctx->pending_n = 0; ctx->pending_direct = false;
	}
	return GAI_OK;
}

int gai_set_launch_config(gai_ctx* ctx, int threads_per_block, int blocks_per_sm)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (threads_per_block) {
		if (threads_per_block < 32 || threads_per_block > 1024 || (threads_per_block & 31)) return fail(ctx, GAI_ERR_INVALID, "threads_per_block must be a multiple of 32 in [32,1024]");
		ctx->fast_threads = threads_per_block;
		ctx->sweep_threads = threads_per_block;
		ctx->threads = threads_per_block > 256 ? 256 : threads_per_block;
	}
	if (blocks_per_sm) {
		if (blocks_per_sm < 1 || blocks_per_sm > 32) return fail(ctx, GAI_ERR_INVALID, "blocks_per_sm must be in [1,32]");
		ctx->blocks_per_sm = blocks_per_sm;
	}
	return GAI_OK;
}

uint64_t gai_launch_count(const gai_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gai_get_launch_config(gai_ctx* ctx, int* threads_per_block, int* plies_per_trip)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (threads_per_block) *threads_per_block = ctx->impl == 2 ? ctx->sweep_threads : ctx->impl == 1 ? ctx->fast_threads : ctx->threads;
	if (plies_per_trip) *plies_per_trip = ctx->impl == 2 ? ctx->sweep_unroll : 1;
	return GAI_OK;
}

int gai_set_rules_impl(gai_ctx* ctx, int impl)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (impl < 0 || impl > 2) return fail(ctx, GAI_ERR_INVALID, "rules impl must be 0 (bitboard baseline), 1 (rotated boards + line LUT, per piece) or 2 (static line sweep)");
	ctx->impl = impl;
	return GAI_OK;
}

// ---- rules sweep ---------------------------------------------------------------------------------
int gai_loa_movegen(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint8_t* counts, gai_loa_move* moves)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (n == 0) return GAI_OK;
	if (!states || !counts || !moves) return fail(ctx, GAI_ERR_INVALID, "gai_loa_movegen: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_loa_movegen");
	const size_t sb = (size_t)n * sizeof(gai_loa_state), mb = (size_t)n * GAI_LOA_MAX_MOVES * sizeof(gai_loa_move);
	EN(ctx->d_in, sb, false); EN(ctx->d_out, n, false); EN(ctx->d_out2, mb, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemsetAsync(ctx->d_out2.p, 0, mb, ctx->stream));
	if (ctx->impl == 2) launch_loa_movegen_sweep(ctx->d_lut, ctx->d_in.p, n, (uint8_t*)ctx->d_out.p, (uint8_t*)ctx->d_out2.p, ctx->prop.multiProcessorCount, ctx->stream);
	else if (ctx->impl == 1) launch_loa_movegen_fast(ctx->d_lut, ctx->d_in.p, n, (uint8_t*)ctx->d_out.p, (uint8_t*)ctx->d_out2.p, ctx->prop.multiProcessorCount, ctx->stream);
	else launch_loa_movegen(ctx->d_in.p, n, (uint8_t*)ctx->d_out.p, (uint8_t*)ctx->d_out2.p, ctx->stream);
	LAUNCHED("k_movegen");
	CU(cudaMemcpyAsync(counts, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(moves, ctx->d_out2.p, mb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

int gai_loa_apply(gai_ctx* ctx, const gai_loa_state* states, const gai_loa_move* mv, uint32_t n, gai_loa_state* out)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (n == 0) return GAI_OK;
	if (!states || !mv || !out) return fail(ctx, GAI_ERR_INVALID, "gai_loa_apply: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_loa_apply");
	const size_t sb = (size_t)n * sizeof(gai_loa_state);
	EN(ctx->d_in, sb, false); EN(ctx->d_in2, (size_t)n * 2, false); EN(ctx->d_out, sb, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemcpyAsync(ctx->d_in2.p, mv, (size_t)n * 2, cudaMemcpyHostToDevice, ctx->stream));
	launch_loa_apply(ctx->d_in.p, (const uint8_t*)ctx->d_in2.p, n, ctx->d_out.p, ctx->stream);
	LAUNCHED("k_apply");
	CU(cudaMemcpyAsync(out, ctx->d_out.p, sb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

int gai_loa_terminal(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint8_t* terminal, uint8_t* score)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (n == 0) return GAI_OK;
	if (!states || !terminal || !score) return fail(ctx, GAI_ERR_INVALID, "gai_loa_terminal: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_loa_terminal");
	const size_t sb = (size_t)n * sizeof(gai_loa_state);
	EN(ctx->d_in, sb, false); EN(ctx->d_out, n, false); EN(ctx->d_out2, (size_t)n * 2, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	launch_loa_terminal(ctx->d_in.p, n, (uint8_t*)ctx->d_out.p, (uint8_t*)ctx->d_out2.p, ctx->stream);
	LAUNCHED("k_terminal");
	CU(cudaMemcpyAsync(terminal, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(score, ctx->d_out2.p, (size_t)n * 2, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

int gai_loa_successor_hash(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint64_t* hash)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (n == 0) return GAI_OK;
	if (!states || !hash) return fail(ctx, GAI_ERR_INVALID, "gai_loa_successor_hash: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_loa_successor_hash");
	const size_t sb = (size_t)n * sizeof(gai_loa_state);
	EN(ctx->d_in, sb, false); EN(ctx->d_out, (size_t)n * 8, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	uint32_t bad = 0;
	if (ctx->impl >= 1) {
		EN(ctx->d_misc, 64, false);
		CU(cudaMemsetAsync((char*)ctx->d_misc.p + 32, 0, 4, ctx->stream));
		launch_loa_successor_hash_fast(ctx->d_lut, ctx->d_in.p, n, (uint64_t*)ctx->d_out.p, (uint32_t*)((char*)ctx->d_misc.p + 32), ctx->prop.multiProcessorCount, ctx->stream);
	} else launch_loa_successor_hash(ctx->d_in.p, n, (uint64_t*)ctx->d_out.p, ctx->stream);
	LAUNCHED("k_successor_hash");
	CU(cudaMemcpyAsync(hash, ctx->d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
	if (ctx->impl >= 1) CU(cudaMemcpyAsync(&bad, (char*)ctx->d_misc.p + 32, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	if (bad) return fail(ctx, GAI_ERR_CUDA, "rotated boards diverged from the plain bitboard after apply (internal consistency check)");
	return GAI_OK;
}

// ---- rollouts --------------------------------------------------------------------------------------
int gai_loa_pack_states(const gai_loa_state* states, uint32_t n, void* boards16, uint32_t* stm_bits)
{
	if (n && (!states || !boards16 || !stm_bits)) return GAI_ERR_INVALID;
	uint64_t* b = (uint64_t*)boards16;
	const uint32_t words = (n + 31u) / 32u;
	for (uint32_t w = 0; w < words; ++w) {
		uint32_t bits = 0;
		const uint32_t lo = w * 32u, hi = lo + 32u < n ? lo + 32u : n;
		for (uint32_t i = lo; i < hi; ++i) {
			b[2 * (size_t)i] = states[i].white; b[2 * (size_t)i + 1] = states[i].black;
			bits |= (uint32_t)(states[i].current_player & 1u) << (i - lo);
		}
		stm_bits[w] = bits;
	}
	return GAI_OK;
}

static int rollout_host_enqueue(gai_ctx* ctx, const gai_loa_state* leaves, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies,
                                uint64_t key, uint64_t first_id, gai_loa_result* out, bool* direct)
{
	if (!leaves) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout: leaves is NULL");
	if (ppl == 0) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout: playouts_per_leaf must be >= 1");
	if (ctx->pending_out) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout: a batch is already in flight on this context (call gai_sync)");
	if (tickets_out(ctx)) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout: ticketed batches are outstanding on this context (gai_wait them first)");
	CU(cudaSetDevice(ctx->device));
	const size_t sb = (size_t)n_leaves * sizeof(gai_loa_state);
	const size_t bb = (size_t)n_leaves * 16, wb = (size_t)((n_leaves + 31u) / 32u) * 4, rb = (size_t)n_leaves * sizeof(gai_loa_result);
	// results: DMA straight into the caller's buffer when it is pinned, else into pinned staging (copied out by gai_sync)
	cudaPointerAttributes po;
	*direct = cudaPointerGetAttributes(&po, out) == cudaSuccess && po.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!*direct) EN(ctx->h_out, rb, true);
	EN(ctx->d_in, sb, false); EN(ctx->d_boards, bb, false); EN(ctx->d_stm, wb, false); EN(ctx->d_out, rb, false);
	// The reference's 24-byte GameState records go to the device as they are and are split into 16-byte boards + side
	// bits THERE (k_split_states): no host-side repack pass.  Pinned caller memory is DMA-ed directly; pageable memory
	// is first copied into pinned staging so that the call may return before the transfer ran (async variant).
	cudaPointerAttributes pa; const void* src = leaves;
	const bool pinned = cudaPointerGetAttributes(&pa, leaves) == cudaSuccess && pa.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!pinned) { EN(ctx->h_in, sb, true); memcpy(ctx->h_in.p, leaves, sb); src = ctx->h_in.p; }
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	CU(cudaMemcpyAsync(ctx->d_in.p, src, sb, cudaMemcpyHostToDevice, ctx->stream));
	launch_loa_split_states(ctx->d_in.p, n_leaves, ctx->d_boards.p, (uint32_t*)ctx->d_stm.p, ctx->stream);
	LAUNCHED("k_split_states");
	int r = enqueue_rollout(ctx, ctx->d_boards.p, (const uint32_t*)ctx->d_stm.p, n_leaves, ppl, max_plies, key, first_id, ctx->d_out.p);
	if (r) return r;
	CU(cudaMemcpyAsync(*direct ? (void*)out : ctx->h_out.p, ctx->d_out.p, rb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaEventRecord(ctx->ev[3], ctx->stream));
	return GAI_OK;
}

int gai_loa_rollout_async(gai_ctx* ctx, const gai_loa_state* leaves, uint32_t n_leaves, uint32_t playouts_per_leaf,
                          uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id, gai_loa_result* out)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (n_leaves == 0) return GAI_OK;
	if (!out) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout_async: out is NULL");
	bool direct = false;
	int r = rollout_host_enqueue(ctx, leaves, n_leaves, playouts_per_leaf, max_plies, philox_key, first_playout_id, out, &direct);
	if (r) return r;
	ctx->pending_out = out; ctx->pending_n = n_leaves; ctx->pending_direct = direct;
	return GAI_OK;
}

int gai_last_rollout_stats(gai_ctx* ctx, gai_rollout_stats* stats)
{
	if (!ctx || !stats) return GAI_ERR_INVALID;
	if (ctx->pending_out) return fail(ctx, GAI_ERR_INVALID, "gai_last_rollout_stats: a batch is still in flight (call gai_sync first)");
	memset(stats, 0, sizeof *stats);
	if (!ctx->d_misc.p) return GAI_OK;
	CU(cudaSetDevice(ctx->device));
	return fill_stats(ctx, stats, true);
}

int gai_loa_rollout(gai_ctx* ctx, const gai_loa_state* leaves, uint32_t n_leaves, uint32_t playouts_per_leaf,
                    uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id,
                    gai_loa_result* out, gai_rollout_stats* stats)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (stats) memset(stats, 0, sizeof *stats);
	if (n_leaves == 0) return GAI_OK;
	int r = gai_loa_rollout_async(ctx, leaves, n_leaves, playouts_per_leaf, max_plies, philox_key, first_playout_id, out);
	if (r) return r;
	r = gai_sync(ctx);
	if (r) return r;
	return fill_stats(ctx, stats, true);
}

// ---- ticketed batches ------------------------------------------------------------------------------
static int submit_batch(gai_ctx* ctx, const gai_loa_state* states, uint32_t n_states, const uint32_t* offsets, uint32_t ppl, uint32_t max_plies,
                        uint64_t key, uint64_t first_id, gai_loa_result* out, int* ticket, const char* who)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (!ticket) return fail(ctx, GAI_ERR_INVALID, "%s: ticket is NULL", who);
	*ticket = -1;
	if (n_states && (!states || !out)) return fail(ctx, GAI_ERR_INVALID, "%s: NULL buffer", who);
	if (ppl == 0) return fail(ctx, GAI_ERR_INVALID, "%s: playouts_per_leaf must be >= 1", who);
	if (ctx->pending_out) return fail(ctx, GAI_ERR_INVALID, "%s: a gai_loa_rollout_async batch is in flight on this context (call gai_sync)", who);
	int si = -1;
	for (int i = 0; i < GAI_MAX_INFLIGHT; ++i) if (!ctx->slots[i].busy) { si = i; break; }
	if (si < 0) return fail(ctx, GAI_ERR_INVALID, "%s: all batch slots of this context are busy (gai_wait one first)", who);
	Slot& sl = ctx->slots[si];
	CU(cudaSetDevice(ctx->device));
	uint32_t n_leaves = n_states;
	if (offsets && n_states) {
		if (offsets[0] != 0) return fail(ctx, GAI_ERR_INVALID, "%s: child_offsets[0] must be 0", who);
		for (uint32_t i = 0; i < n_states; ++i)
			if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > GAI_LOA_MAX_MOVES) return fail(ctx, GAI_ERR_INVALID, "%s: child_offsets must be non-decreasing with at most 96 children per parent", who);
		n_leaves = offsets[n_states];
	}
	sl.children = offsets != nullptr; sl.out = out; sl.n_out = n_leaves; sl.direct = false;
	EN(sl.h_stat, 64, true); EN(sl.d_misc, 128, false);
	memset(sl.h_stat.p, 0, 64);
	CU(cudaEventRecord(sl.ev[0], ctx->stream));
	if (n_leaves == 0) {                       // nothing to play: the ticket completes immediately
		CU(cudaEventRecord(sl.ev[1], ctx->stream)); CU(cudaEventRecord(sl.ev[2], ctx->stream)); CU(cudaEventRecord(sl.ev[3], ctx->stream));
		sl.busy = true; sl.gen += 1; *ticket = (int)(sl.gen * GAI_MAX_INFLIGHT + (unsigned)si);
		return GAI_OK;
	}
	const size_t sb = (size_t)n_states * sizeof(gai_loa_state), ob = offsets ? (size_t)(n_states + 1) * 4 : 0;
	const size_t bb = (size_t)n_leaves * 16, wb = (size_t)((n_leaves + 31u) / 32u) * 4, rb = (size_t)n_leaves * sizeof(gai_loa_result);
	cudaPointerAttributes po;
	sl.direct = cudaPointerGetAttributes(&po, out) == cudaSuccess && po.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!sl.direct) EN(sl.h_out, rb, true);
	EN(sl.d_in, sb, false); EN(sl.d_boards, bb, false); EN(sl.d_stm, wb, false); EN(sl.d_out, rb, false);
	cudaPointerAttributes pa; const void* src = states;
	const bool pinned = cudaPointerGetAttributes(&pa, states) == cudaSuccess && pa.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!pinned) { EN(sl.h_in, sb, true); memcpy(sl.h_in.p, states, sb); src = sl.h_in.p; }
	CU(cudaMemcpyAsync(sl.d_in.p, src, sb, cudaMemcpyHostToDevice, ctx->stream));
	uint32_t* d_bad = (uint32_t*)((char*)sl.d_misc.p + 64);
	CU(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
	if (offsets) {
		EN(sl.h_off, ob, true); EN(sl.d_off, ob, false); EN(sl.d_side, n_leaves, false);
		memcpy(sl.h_off.p, offsets, ob);
		CU(cudaMemcpyAsync(sl.d_off.p, sl.h_off.p, ob, cudaMemcpyHostToDevice, ctx->stream));
		launch_loa_expand_children(sl.d_in.p, n_states, (const uint32_t*)sl.d_off.p, sl.d_boards.p, (uint8_t*)sl.d_side.p, d_bad, ctx->stream);
		LAUNCHED("k_expand_children");
		launch_loa_pack_stm((const uint8_t*)sl.d_side.p, n_leaves, (uint32_t*)sl.d_stm.p, ctx->stream);
		LAUNCHED("k_pack_stm");
	} else {
		launch_loa_split_states(sl.d_in.p, n_leaves, sl.d_boards.p, (uint32_t*)sl.d_stm.p, ctx->stream);
		LAUNCHED("k_split_states");
	}
	int r = enqueue_rollout(ctx, sl.d_misc, sl.d_prep, sl.ev[1], sl.ev[2], sl.d_boards.p, (const uint32_t*)sl.d_stm.p, n_leaves, ppl, max_plies, key, first_id, sl.d_out.p);
	if (r) return r;
	CU(cudaMemcpyAsync(sl.direct ? (void*)out : sl.h_out.p, sl.d_out.p, rb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(sl.h_stat.p, sl.d_misc.p, 16, cudaMemcpyDeviceToHost, ctx->stream));                   // {playouts, plies}
	CU(cudaMemcpyAsync((char*)sl.h_stat.p + 16, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaEventRecord(sl.ev[3], ctx->stream));
	sl.busy = true; sl.gen += 1; *ticket = (int)(sl.gen * GAI_MAX_INFLIGHT + (unsigned)si);
	return GAI_OK;
}

int gai_loa_rollout_submit(gai_ctx* ctx, const gai_loa_state* leaves, uint32_t n_leaves, uint32_t playouts_per_leaf,
                           uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id, gai_loa_result* out, int* ticket)
{
	return submit_batch(ctx, leaves, n_leaves, nullptr, playouts_per_leaf, max_plies, philox_key, first_playout_id, out, ticket, "gai_loa_rollout_submit");
}

int gai_loa_rollout_children_submit(gai_ctx* ctx, const gai_loa_state* parents, uint32_t n_parents, const uint32_t* child_offsets,
                                    uint32_t playouts_per_leaf, uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id,
                                    gai_loa_result* out, int* ticket)
{
	if (ctx && !child_offsets) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout_children_submit: child_offsets is NULL");
	return submit_batch(ctx, parents, n_parents, child_offsets, playouts_per_leaf, max_plies, philox_key, first_playout_id, out, ticket, "gai_loa_rollout_children_submit");
}

int gai_inflight(const gai_ctx* ctx) { return ctx ? tickets_out(ctx) : 0; }

int gai_wait(gai_ctx* ctx, int ticket, gai_rollout_stats* stats)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (stats) memset(stats, 0, sizeof *stats);
	if (ticket < 0) return fail(ctx, GAI_ERR_INVALID, "gai_wait: bad ticket");
	Slot& sl = ctx->slots[ticket % GAI_MAX_INFLIGHT];
	if (!sl.busy || sl.gen != (unsigned)ticket / GAI_MAX_INFLIGHT) return fail(ctx, GAI_ERR_INVALID, "gai_wait: no such ticket outstanding");
	sl.busy = false;                              // the slot is free again whatever happens below
	CU(cudaSetDevice(ctx->device));
	CU(cudaEventSynchronize(sl.ev[3]));
	const uint64_t* hs = (const uint64_t*)sl.h_stat.p;
	if (sl.n_out && !sl.direct) memcpy(sl.out, sl.h_out.p, (size_t)sl.n_out * sizeof(gai_loa_result));
	if (stats) {
		stats->playouts = hs[0]; stats->plies = hs[1];
		if (sl.n_out) { CU(cudaEventElapsedTime(&stats->kernel_ms, sl.ev[1], sl.ev[2])); CU(cudaEventElapsedTime(&stats->total_ms, sl.ev[0], sl.ev[3])); }
	}
	if (((const uint32_t*)sl.h_stat.p)[4]) return fail(ctx, GAI_ERR_INVALID, "gai_wait: a parent's move count on the device differs from child_offsets (caller and device rules disagree)");
	return GAI_OK;
}

int gai_loa_rollout_device(gai_ctx* ctx, const void* d_boards, const uint32_t* d_stm, uint32_t n_leaves,
                           uint32_t playouts_per_leaf, uint32_t max_plies, uint64_t philox_key,
                           uint64_t first_playout_id, void* d_out, gai_rollout_stats* stats)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (stats) memset(stats, 0, sizeof *stats);
	if (n_leaves == 0) return GAI_OK;
	if (!d_boards || !d_stm || !d_out) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout_device: NULL buffer");
	if (playouts_per_leaf == 0) return fail(ctx, GAI_ERR_INVALID, "gai_loa_rollout_device: playouts_per_leaf must be >= 1");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_loa_rollout_device");
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	int r = enqueue_rollout(ctx, d_boards, d_stm, n_leaves, playouts_per_leaf, max_plies, philox_key, first_playout_id, d_out);
	if (r) return r;
	CU(cudaEventRecord(ctx->ev[3], ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return fill_stats(ctx, stats, true);
}

int gai_loa_playout_trace(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint32_t max_plies,
                          uint64_t philox_key, uint64_t first_playout_id,
                          uint8_t* outcome, uint32_t* plies, gai_loa_state* final_states)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (n == 0) return GAI_OK;
	if (!states || !outcome || !plies || !final_states) return fail(ctx, GAI_ERR_INVALID, "gai_loa_playout_trace: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_loa_playout_trace");
	const size_t sb = (size_t)n * sizeof(gai_loa_state);
	EN(ctx->d_in, sb, false); EN(ctx->d_out, n, false); EN(ctx->d_out2, (size_t)n * 4, false); EN(ctx->d_out3, sb, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	if (ctx->impl == 2)          // the static-sweep step itself (sweep + select + apply with chi + chi-gated flood fill)
		launch_loa_playout_trace_sweep(ctx->d_lut, ctx->d_in.p, n, max_plies, philox_key, first_playout_id, (uint8_t*)ctx->d_out.p,
		                               (uint32_t*)ctx->d_out2.p, ctx->d_out3.p, ctx->prop.multiProcessorCount, ctx->stream);
	else if (ctx->impl == 1)
		launch_loa_playout_trace_fast(ctx->d_lut, ctx->d_in.p, n, max_plies, philox_key, first_playout_id, (uint8_t*)ctx->d_out.p,
		                              (uint32_t*)ctx->d_out2.p, ctx->d_out3.p, ctx->prop.multiProcessorCount, ctx->stream);
	else
		launch_loa_playout_trace(ctx->d_in.p, n, max_plies, philox_key, first_playout_id, (uint8_t*)ctx->d_out.p,
		                         (uint32_t*)ctx->d_out2.p, ctx->d_out3.p, ctx->stream);
	LAUNCHED("k_playout_trace");
	CU(cudaMemcpyAsync(outcome, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(plies, ctx->d_out2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(final_states, ctx->d_out3.p, sb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

int gai_loa_gen_reachable(gai_ctx* ctx, uint32_t n, uint64_t key, uint32_t max_k,
                          void* d_boards, uint32_t* d_stm, gai_loa_state* host_states)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (n == 0) return GAI_OK;
	if (max_k == 0) return fail(ctx, GAI_ERR_INVALID, "gai_loa_gen_reachable: max_k must be >= 1");
	if ((d_boards == nullptr) != (d_stm == nullptr)) return fail(ctx, GAI_ERR_INVALID, "gai_loa_gen_reachable: d_boards and d_stm go together");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_loa_gen_reachable");
	const size_t sb = (size_t)n * sizeof(gai_loa_state);
	EN(ctx->d_in2, n, false);
	if (host_states) EN(ctx->d_out3, sb, false);
	launch_loa_gen_reachable(n, key, max_k, d_boards, host_states ? ctx->d_out3.p : nullptr, (uint8_t*)ctx->d_in2.p, ctx->stream);
	LAUNCHED("k_gen_reachable");
	if (d_stm) { launch_loa_pack_stm((const uint8_t*)ctx->d_in2.p, n, d_stm, ctx->stream); LAUNCHED("k_pack_stm"); }
	if (host_states) CU(cudaMemcpyAsync(host_states, ctx->d_out3.p, sb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

// ---- Pan ---------------------------------------------------------------------------------------------
static inline uint64_t known_hand_count(uint64_t w) { return w & 15; }
static int pan_check_np(gai_ctx* ctx, int np)
{
	if (np < 2 || np > 4) return fail(ctx, GAI_ERR_INVALID, "num_players must be 2, 3 or 4");
	return GAI_OK;
}
static int pan_bad(gai_ctx* ctx, uint32_t bad)
{
	if (bad) return fail(ctx, GAI_ERR_INVALID, "Pan state with fuzzy (01/10) cards: the GPU path needs a determinized deal (every card 00 or 11)");
	return GAI_OK;
}
#define PAN_BAD_PTR ((uint32_t*)((char*)ctx->d_misc.p + 48))

int gai_pan_movegen(gai_ctx* ctx, const gai_pan_state* states, uint32_t n, int np, uint8_t* counts, gai_pan_move* moves)
{
	if (!ctx) return GAI_ERR_INVALID;
	int r = pan_check_np(ctx, np); if (r) return r;
	if (n == 0) return GAI_OK;
	if (!states || !counts || !moves) return fail(ctx, GAI_ERR_INVALID, "gai_pan_movegen: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_pan_movegen");
	const size_t sb = (size_t)n * 40, mb = (size_t)n * GAI_PAN_MAX_MOVES * 8;
	EN(ctx->d_in, sb, false); EN(ctx->d_out, n, false); EN(ctx->d_out2, mb, false); EN(ctx->d_misc, 64, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemsetAsync(ctx->d_out2.p, 0, mb, ctx->stream));
	CU(cudaMemsetAsync(PAN_BAD_PTR, 0, 4, ctx->stream));
	launch_pan_movegen(ctx->d_in.p, n, np, (uint8_t*)ctx->d_out.p, (uint64_t*)ctx->d_out2.p, PAN_BAD_PTR, ctx->stream);
	LAUNCHED("k_pan_movegen");
	uint32_t bad = 0;
	CU(cudaMemcpyAsync(counts, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(moves, ctx->d_out2.p, mb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(&bad, PAN_BAD_PTR, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return pan_bad(ctx, bad);
}

int gai_pan_apply(gai_ctx* ctx, const gai_pan_state* states, const gai_pan_move* mv, uint32_t n, int np, gai_pan_state* out)
{
	if (!ctx) return GAI_ERR_INVALID;
	int r = pan_check_np(ctx, np); if (r) return r;
	if (n == 0) return GAI_OK;
	if (!states || !mv || !out) return fail(ctx, GAI_ERR_INVALID, "gai_pan_apply: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_pan_apply");
	const size_t sb = (size_t)n * 40;
	EN(ctx->d_in, sb, false); EN(ctx->d_in2, (size_t)n * 8, false); EN(ctx->d_out, sb, false); EN(ctx->d_misc, 64, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemcpyAsync(ctx->d_in2.p, mv, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemsetAsync(PAN_BAD_PTR, 0, 4, ctx->stream));
	launch_pan_apply(ctx->d_in.p, (const uint64_t*)ctx->d_in2.p, n, np, ctx->d_out.p, PAN_BAD_PTR, ctx->stream);
	LAUNCHED("k_pan_apply");
	uint32_t bad = 0;
	CU(cudaMemcpyAsync(out, ctx->d_out.p, sb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(&bad, PAN_BAD_PTR, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return pan_bad(ctx, bad);
}

int gai_pan_terminal(gai_ctx* ctx, const gai_pan_state* states, uint32_t n, int np, uint8_t* terminal,
                     int32_t* score, int32_t* ev0, int32_t* ev1)
{
	if (!ctx) return GAI_ERR_INVALID;
	int r = pan_check_np(ctx, np); if (r) return r;
	if (n == 0) return GAI_OK;
	if (!states || !terminal || !score || !ev0 || !ev1) return fail(ctx, GAI_ERR_INVALID, "gai_pan_terminal: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_pan_terminal");
	const size_t sb = (size_t)n * 40, vb = (size_t)n * 16;
	EN(ctx->d_in, sb, false); EN(ctx->d_out, n, false); EN(ctx->d_out2, 3 * vb, false); EN(ctx->d_misc, 64, false);
	int* d_v = (int*)ctx->d_out2.p;
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemsetAsync(PAN_BAD_PTR, 0, 4, ctx->stream));
	launch_pan_terminal(ctx->d_in.p, n, np, (uint8_t*)ctx->d_out.p, d_v, d_v + 4 * (size_t)n, d_v + 8 * (size_t)n, PAN_BAD_PTR, ctx->stream);
	LAUNCHED("k_pan_terminal");
	uint32_t bad = 0;
	CU(cudaMemcpyAsync(terminal, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(score, d_v, vb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(ev0, d_v + 4 * (size_t)n, vb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(ev1, d_v + 8 * (size_t)n, vb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(&bad, PAN_BAD_PTR, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return pan_bad(ctx, bad);
}

// Host buffers in, host buffers out, in pipeline stages of GAI_PAN_CHUNK deals: stage c+1 is uploaded (own copy stream) while
// stage c plays out (compute stream) and stage c-1 is downloaded (second copy stream); a pageable buffer is staged through pinned
// memory one stage at a time, a PINNED caller buffer is DMA-ed in place.  Deals are independent, a playout's identity is
// (key, first_id + leaf * ppl + j): the stages change nothing in the results.  stats->kernel_ms = the stages' kernel time (all
// kernels run back to back on the one compute stream), total_ms = first upload to last download.
int gai_pan_rollout(gai_ctx* ctx, const gai_pan_state* leaves, uint32_t n_leaves, int np, uint32_t ppl,
                    uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id, gai_pan_result* out, gai_rollout_stats* stats)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (stats) memset(stats, 0, sizeof *stats);
	int r = pan_check_np(ctx, np); if (r) return r;
	if (n_leaves == 0) return GAI_OK;
	if (!leaves || !out) return fail(ctx, GAI_ERR_INVALID, "gai_pan_rollout: NULL buffer");
	if (ppl == 0) return fail(ctx, GAI_ERR_INVALID, "gai_pan_rollout: playouts_per_leaf must be >= 1");
	if (eval_kind != 0 && eval_kind != 1) return fail(ctx, GAI_ERR_INVALID, "gai_pan_rollout: eval_kind must be 0 (num_cards) or 1 (num_cards_weighted)");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_pan_rollout");
	if (!ctx->pan_in) {
		CU(cudaStreamCreateWithFlags(&ctx->pan_in, cudaStreamNonBlocking));
		CU(cudaStreamCreateWithFlags(&ctx->pan_out, cudaStreamNonBlocking));
		for (int k = 0; k < 4; ++k) for (int c = 0; c < GAI_PAN_MAX_CHUNKS; ++c)
			CU(cudaEventCreateWithFlags(&ctx->pan_ev[k][c], (k == 1 || k == 2) ? cudaEventDefault : cudaEventDisableTiming));
	}
	uint32_t chunk = GAI_PAN_CHUNK;
	if ((n_leaves + chunk - 1) / chunk > (uint32_t)GAI_PAN_MAX_CHUNKS) chunk = (n_leaves + GAI_PAN_MAX_CHUNKS - 1) / GAI_PAN_MAX_CHUNKS;
	const int nch = (int)((n_leaves + chunk - 1) / chunk);
	const size_t sb = (size_t)n_leaves * 40, rb = (size_t)n_leaves * sizeof(gai_pan_result);
	cudaPointerAttributes pa{}, po{};
	const bool pin_in = cudaPointerGetAttributes(&pa, leaves) == cudaSuccess && pa.type == cudaMemoryTypeHost;
	const bool pin_out = cudaPointerGetAttributes(&po, out) == cudaSuccess && po.type == cudaMemoryTypeHost;
	cudaGetLastError();                                   // a pageable pointer leaves an error behind on old drivers
	if (!pin_in) EN(ctx->h_in, sb, true);
	if (!pin_out) EN(ctx->h_out, rb, true);
	EN(ctx->d_in, sb, false); EN(ctx->d_out, rb, false); EN(ctx->d_prep, (size_t)n_leaves * 32, false);
	EN(ctx->pan_misc, 64 * (size_t)(GAI_PAN_MAX_CHUNKS + 1), false);
	uint32_t* d_bad = (uint32_t*)((char*)ctx->pan_misc.p + 64 * (size_t)GAI_PAN_MAX_CHUNKS);
	const char* src = pin_in ? (const char*)leaves : (const char*)ctx->h_in.p;
	char* dst = pin_out ? (char*)out : (char*)ctx->h_out.p;
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	CU(cudaStreamWaitEvent(ctx->pan_in, ctx->ev[0], 0));
	CU(cudaMemsetAsync(ctx->pan_misc.p, 0, 64 * (size_t)(GAI_PAN_MAX_CHUNKS + 1), ctx->stream));
	CU(cudaMemsetAsync(ctx->d_out.p, 0, rb, ctx->stream));
	for (int c = 0; c < nch; ++c) {
		const uint32_t off = (uint32_t)c * chunk, nc = std::min(chunk, n_leaves - off);
		if (!pin_in) memcpy((char*)ctx->h_in.p + (size_t)off * 40, (const char*)leaves + (size_t)off * 40, (size_t)nc * 40);
		CU(cudaMemcpyAsync((char*)ctx->d_in.p + (size_t)off * 40, src + (size_t)off * 40, (size_t)nc * 40, cudaMemcpyHostToDevice, ctx->pan_in));
		CU(cudaEventRecord(ctx->pan_ev[0][c], ctx->pan_in));
		CU(cudaStreamWaitEvent(ctx->stream, ctx->pan_ev[0][c], 0));
		CU(cudaEventRecord(ctx->pan_ev[1][c], ctx->stream));
		const uint64_t total = (uint64_t)nc * ppl;
		int blocks = ctx->prop.multiProcessorCount * 8;
		if ((total + 255) / 256 < (uint64_t)blocks) blocks = (int)((total + 255) / 256);
		uint64_t* d_st = (uint64_t*)((char*)ctx->pan_misc.p + 64 * (size_t)c);
		launch_pan_rollout((char*)ctx->d_in.p + (size_t)off * 40, (char*)ctx->d_prep.p + (size_t)off * 32, nc, np, ppl, max_plies, eval_kind, key,
		                   first_id + (uint64_t)off * ppl, (uint32_t*)((char*)ctx->d_out.p + (size_t)off * sizeof(gai_pan_result)),
		                   d_st, (unsigned long long*)d_st + 2, d_bad, blocks, ctx->stream);
		LAUNCHED("k_pan_prepare"); LAUNCHED("k_pan_rollout");
		CU(cudaEventRecord(ctx->pan_ev[2][c], ctx->stream));
		CU(cudaStreamWaitEvent(ctx->pan_out, ctx->pan_ev[2][c], 0));
		CU(cudaMemcpyAsync(dst + (size_t)off * sizeof(gai_pan_result), (char*)ctx->d_out.p + (size_t)off * sizeof(gai_pan_result),
		                   (size_t)nc * sizeof(gai_pan_result), cudaMemcpyDeviceToHost, ctx->pan_out));
		CU(cudaEventRecord(ctx->pan_ev[3][c], ctx->pan_out));
	}
	CU(cudaStreamWaitEvent(ctx->stream, ctx->pan_ev[3][nch - 1], 0));
	CU(cudaEventRecord(ctx->ev[3], ctx->stream));
	for (int c = 0; c < nch; ++c) {                        // a pageable `out` is filled stage by stage while later stages still run
		CU(cudaEventSynchronize(ctx->pan_ev[3][c]));
		const uint32_t off = (uint32_t)c * chunk, nc = std::min(chunk, n_leaves - off);
		if (!pin_out) memcpy((char*)out + (size_t)off * sizeof(gai_pan_result), (char*)ctx->h_out.p + (size_t)off * sizeof(gai_pan_result), (size_t)nc * sizeof(gai_pan_result));
	}
	CU(cudaStreamSynchronize(ctx->stream));
	uint64_t h[8 * GAI_PAN_MAX_CHUNKS + 8];
	CU(cudaMemcpy(h, ctx->pan_misc.p, 64 * (size_t)(GAI_PAN_MAX_CHUNKS + 1), cudaMemcpyDeviceToHost));
	r = pan_bad(ctx, (uint32_t)h[8 * GAI_PAN_MAX_CHUNKS]); if (r) return r;
	if (stats) {
		for (int c = 0; c < nch; ++c) {
			stats->playouts += h[8 * c]; stats->plies += h[8 * c + 1];
			float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->pan_ev[1][c], ctx->pan_ev[2][c])); stats->kernel_ms += ms;
		}
		CU(cudaEventElapsedTime(&stats->total_ms, ctx->ev[0], ctx->ev[3]));
	}
	return GAI_OK;
}

int gai_pan_rollout_device(gai_ctx* ctx, const void* d_states, uint32_t n_leaves, int np, uint32_t ppl, uint32_t max_plies, int eval_kind,
                           uint64_t key, uint64_t first_id, void* d_out, gai_rollout_stats* stats)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (stats) memset(stats, 0, sizeof *stats);
	int r = pan_check_np(ctx, np); if (r) return r;
	if (n_leaves == 0) return GAI_OK;
	if (!d_states || !d_out) return fail(ctx, GAI_ERR_INVALID, "gai_pan_rollout_device: NULL buffer");
	if (ppl == 0) return fail(ctx, GAI_ERR_INVALID, "gai_pan_rollout_device: playouts_per_leaf must be >= 1");
	if (eval_kind != 0 && eval_kind != 1) return fail(ctx, GAI_ERR_INVALID, "gai_pan_rollout_device: eval_kind must be 0 (num_cards) or 1 (num_cards_weighted)");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_pan_rollout_device");
	EN(ctx->d_misc, 64, false); EN(ctx->d_prep, (size_t)n_leaves * 32, false);
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	CU(cudaMemsetAsync(ctx->d_misc.p, 0, 64, ctx->stream));
	CU(cudaMemsetAsync(d_out, 0, (size_t)n_leaves * sizeof(gai_pan_result), ctx->stream));
	const uint64_t total = (uint64_t)n_leaves * ppl;
	int blocks = ctx->prop.multiProcessorCount * 8;
	if ((total + 255) / 256 < (uint64_t)blocks) blocks = (int)((total + 255) / 256);
	CU(cudaEventRecord(ctx->ev[1], ctx->stream));
	launch_pan_rollout(d_states, ctx->d_prep.p, n_leaves, np, ppl, max_plies, eval_kind, key, first_id, (uint32_t*)d_out,
	                   (uint64_t*)ctx->d_misc.p, (unsigned long long*)ctx->d_misc.p + 2, PAN_BAD_PTR, blocks, ctx->stream);
	LAUNCHED("k_pan_prepare"); LAUNCHED("k_pan_rollout");
	CU(cudaEventRecord(ctx->ev[2], ctx->stream));
	uint32_t bad = 0;
	CU(cudaMemcpyAsync(&bad, PAN_BAD_PTR, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaEventRecord(ctx->ev[3], ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	r = pan_bad(ctx, bad); if (r) return r;
	return fill_stats(ctx, stats, true);
}

int gai_pan_playout_trace(gai_ctx* ctx, const gai_pan_state* states, uint32_t n, int np, uint32_t max_plies, int eval_kind,
                          uint64_t key, uint64_t first_id, uint8_t* code, uint32_t* plies, int32_t* value, gai_pan_state* final_states)
{
	if (!ctx) return GAI_ERR_INVALID;
	int r = pan_check_np(ctx, np); if (r) return r;
	if (n == 0) return GAI_OK;
	if (!states || !code || !plies || !value || !final_states) return fail(ctx, GAI_ERR_INVALID, "gai_pan_playout_trace: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_pan_playout_trace");
	const size_t sb = (size_t)n * 40;
	EN(ctx->d_in, sb, false); EN(ctx->d_out, n, false); EN(ctx->d_out2, (size_t)n * 20, false); EN(ctx->d_out3, sb, false); EN(ctx->d_misc, 64, false);
	uint32_t* d_pl = (uint32_t*)ctx->d_out2.p; int* d_val = (int*)(d_pl + n);
	CU(cudaMemcpyAsync(ctx->d_in.p, states, sb, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemsetAsync(PAN_BAD_PTR, 0, 4, ctx->stream));
	launch_pan_playout_trace(ctx->d_in.p, n, np, max_plies, eval_kind, key, first_id, (uint8_t*)ctx->d_out.p, d_pl, d_val, ctx->d_out3.p, PAN_BAD_PTR, ctx->stream);
	LAUNCHED("k_pan_playout_trace");
	uint32_t bad = 0;
	CU(cudaMemcpyAsync(code, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(plies, d_pl, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(value, d_val, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(final_states, ctx->d_out3.p, sb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(&bad, PAN_BAD_PTR, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return pan_bad(ctx, bad);
}

int gai_pan_determinize(const gai_pan_state* known, int np, int player, uint32_t n_samples,
                        uint64_t key, uint64_t first_id, gai_pan_state* out)
{
	if (!known || (!out && n_samples) || np < 2 || np > 4 || player < 0 || player >= np) return GAI_ERR_INVALID;
	const uint64_t M48 = 0xFFFFFFFFFFFFull;
	const uint64_t stack = known->w[0] & M48;
	const uint64_t own = (known->w[1 + player] >> 4) & M48;
	// own hand and stack must be certain (00 / 11)
	auto crisp = [](uint64_t c) { return (((c >> 1) ^ c) & 0x555555555555ull) == 0; };
	if (!crisp(stack) || !crisp(own)) return GAI_ERR_INVALID;
	// What the view KNOWS about the other hands: after a take the reference's apply_move does hand.cards |= cards
	// (GraWPanaZasadyV2.cpp:541-545), so those cards are certain (11) in the taker's hand of every player's view; with three
	// or four players an unseen card is only ever marked 10 / 01 (CreatePlayerKnownState :216-233), so 11 in another hand
	// means "held by that player".  (Two players: the view is complete, every unseen card is the opponent's, marked 11.)
	int unseen[24], nu = 0, cnt[4] = { 0, 0, 0, 0 }, need[4] = { 0, 0, 0, 0 };
	uint64_t held[4] = { 0, 0, 0, 0 };
	for (int p = 0; p < np; ++p) cnt[p] = (int)(known_hand_count(known->w[1 + p]));
	int total_need = 0;
	for (int c = 0; c < 24; ++c) {
		if (((stack >> (2 * c)) & 3) || ((own >> (2 * c)) & 3)) continue;
		int holder = -1;
		if (np >= 3) for (int p = 0; p < np; ++p) if (p != player && (((known->w[1 + p] >> 4) >> (2 * c)) & 3) == 3) { holder = p; break; }
		if (holder >= 0) held[holder] |= 3ull << (2 * c); else unseen[nu++] = c;
	}
	for (int p = 0; p < np; ++p) {
		if (p == player) continue;
		need[p] = cnt[p] - __builtin_popcountll(held[p]) / 2;
		if (need[p] < 0) return GAI_ERR_INVALID;            // (a count field wrapped past 15, or an inconsistent state)
		total_need += need[p];
	}
	if (total_need != nu) return GAI_ERR_INVALID;
	// rule constraint: on an empty stack the player to move holds the 9 of hearts (GraWPanaZasadyV2.cpp:127-130,415-418)
	const int cp = (int)((known->w[0] >> 48) & 7);
	const bool pin9h = stack == 0 && cp != player && cp < np && nu > 0 && unseen[0] == 0 && need[cp] > 0;
	for (uint32_t i = 0; i < n_samples; ++i) {
		int deck[24];
		for (int k = 0; k < nu; ++k) deck[k] = unseen[k];
		uint32_t draw = 0; gai::Philox4 blk{};
		const int lo = pin9h ? 1 : 0;                      // deck[0] = 9h stays out of the shuffle and is dealt to cp
		for (int k = nu - 1; k > lo; --k) {                // Fisher-Yates, one choice word per swap
			if ((draw & 3u) == 0) blk = gai::choice_block(key, first_id + i, draw >> 2);
			const uint32_t j = (uint32_t)lo + gai::choice_index(blk.v[draw & 3u], (uint32_t)(k + 1 - lo));
			++draw;
			const int t = deck[k]; deck[k] = deck[j]; deck[j] = t;
		}
		gai_pan_state s; memset(&s, 0, sizeof s);
		s.w[0] = known->w[0];
		int at = pin9h ? 1 : 0;
		for (int p = 0; p < np; ++p) {
			uint64_t cards = 0;
			if (p == player) cards = own;
			else {
				cards = held[p];                           // what the view already knows this player holds
				int take = need[p];
				if (pin9h && p == cp) { cards |= 3ull; --take; }
				for (int k = 0; k < take; ++k) cards |= 3ull << (2 * deck[at++]);
			}
			s.w[1 + p] = (uint64_t)cnt[p] | (cards << 4);
		}
		out[i] = s;
	}
	return GAI_OK;
}

// ---- device memory helpers -------------------------------------------------------------------------
int gai_dev_alloc(gai_ctx* ctx, uint64_t bytes, void** d_ptr)
{
	if (!ctx || !d_ptr) return GAI_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
	if (e != cudaSuccess) return fail(ctx, GAI_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
	return GAI_OK;
}
int gai_dev_free(gai_ctx* ctx, void* d_ptr)
{
	if (!ctx) return GAI_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaFree(d_ptr));
	return GAI_OK;
}
int gai_memcpy_h2d(gai_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes)
{
	if (!ctx) return GAI_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}
int gai_memcpy_d2h(gai_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes)
{
	if (!ctx) return GAI_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

// ---- NCCL (root parallelism) -----------------------------------------------------------------------
static int nccl_load(NcclApi& a, char* err, size_t errlen)
{
	if (a.h) return GAI_OK;
	const char* names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char* nm : names) { a.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (a.h) break; }
	if (!a.h) { snprintf(err, errlen, "dlopen(libnccl.so.2) failed: %s", dlerror()); return GAI_ERR_NCCL; }
	a.GetUniqueId = (int (*)(void*))dlsym(a.h, "ncclGetUniqueId");
	a.CommInitRank = (int (*)(ncclComm_t*, int, Id128, int))dlsym(a.h, "ncclCommInitRank");
	a.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(a.h, "ncclAllReduce");
	a.CommDestroy = (int (*)(ncclComm_t))dlsym(a.h, "ncclCommDestroy");
	a.GetErrorString = (const char* (*)(int))dlsym(a.h, "ncclGetErrorString");
	a.GroupStart = (int (*)())dlsym(a.h, "ncclGroupStart");
	a.GroupEnd = (int (*)())dlsym(a.h, "ncclGroupEnd");
	if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) { snprintf(err, errlen, "libnccl lacks a required symbol"); return GAI_ERR_NCCL; }
	return GAI_OK;
}
#define NC(call) do { int e_ = (call); if (e_ != 0) return fail(ctx, GAI_ERR_NCCL, "%s: %s", #call, ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(e_) : "nccl error"); } while (0)

int gai_comm_unique_id(uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES])
{
	static NcclApi api;
	if (!id) return GAI_ERR_INVALID;
	int r = nccl_load(api, g_create_error, sizeof g_create_error);
	if (r) return r;
	Id128 u; memset(&u, 0, sizeof u);
	if (api.GetUniqueId(&u) != 0) return fail(nullptr, GAI_ERR_NCCL, "ncclGetUniqueId failed");
	memcpy(id, &u, sizeof u);
	return GAI_OK;
}

int gai_comm_init(gai_ctx* ctx, int rank, int world, const uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES])
{
	if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return ctx ? fail(ctx, GAI_ERR_INVALID, "gai_comm_init: bad arguments") : GAI_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int r = nccl_load(ctx->nccl, ctx->err, sizeof ctx->err);
	if (r) return r;
	Id128 u; memcpy(&u, id, sizeof u);
	NC(ctx->nccl.CommInitRank(&ctx->comm, world, u, rank));
	ctx->rank = rank; ctx->world = world;
	return GAI_OK;
}

// ncclDataType_t: ncclUint32 = 3, ncclUint64 = 5, ncclFloat32 = 7 ; ncclRedOp_t: ncclSum = 0 (nccl.h, stable since 2.0)
int gai_allreduce_root(gai_ctx* ctx, uint32_t* visits, float* values, uint32_t count)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (!ctx->comm) return fail(ctx, GAI_ERR_NCCL, "gai_allreduce_root: communicator not initialised (gai_comm_init)");
	if (count == 0) return GAI_OK;
	if (!visits || !values) return fail(ctx, GAI_ERR_INVALID, "gai_allreduce_root: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_allreduce_root");
	const size_t vb = (size_t)count * 4, fb = (size_t)count * 8;
	EN(ctx->d_in, vb + fb, false);
	char* d = (char*)ctx->d_in.p;
	CU(cudaMemcpyAsync(d, visits, vb, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemcpyAsync(d + vb, values, fb, cudaMemcpyHostToDevice, ctx->stream));
	NC(ctx->nccl.GroupStart());
	NC(ctx->nccl.AllReduce(d, d, count, 3, 0, ctx->comm, ctx->stream));
	NC(ctx->nccl.AllReduce(d + vb, d + vb, (size_t)count * 2, 7, 0, ctx->comm, ctx->stream));
	NC(ctx->nccl.GroupEnd());
	CU(cudaMemcpyAsync(visits, d, vb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(values, d + vb, fb, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

int gai_allreduce_u64(gai_ctx* ctx, uint64_t* v, uint32_t count)
{
	if (!ctx) return GAI_ERR_INVALID;
	if (!ctx->comm) return fail(ctx, GAI_ERR_NCCL, "gai_allreduce_u64: communicator not initialised (gai_comm_init)");
	if (count == 0) return GAI_OK;
	if (!v) return fail(ctx, GAI_ERR_INVALID, "gai_allreduce_u64: NULL buffer");
	CU(cudaSetDevice(ctx->device));
	QUIESCE("gai_allreduce_u64");
	EN(ctx->d_in, (size_t)count * 8, false);
	CU(cudaMemcpyAsync(ctx->d_in.p, v, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
	NC(ctx->nccl.AllReduce(ctx->d_in.p, ctx->d_in.p, count, 5, 0, ctx->comm, ctx->stream));
	CU(cudaMemcpyAsync(v, ctx->d_in.p, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return GAI_OK;
}

} // extern "C"
