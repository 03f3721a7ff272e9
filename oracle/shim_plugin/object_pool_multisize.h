// oracle/shim_plugin/object_pool_multisize.h -- TEST INFRASTRUCTURE ONLY; used ONLY for oracle/_ref/librefmctsplayer.so.
//
// Stand-in for the reference's c++/common/utils/object_pool_multisize.h with the same interface (alloc / free / free_all /
// usage counters, MCTSPlayer.h:131-132) on top of the system allocator.  Why: the reference's free_all() sets
// num_free_chunks = 0 instead of the block size (:157-165) and selection_playOut calls it after every simulation
// (MCTSPlayer.cpp:467), so every later alloc opens a fresh 16 384-chunk block and scans an ever longer std::map -- the player
// slows down super-linearly with the number of simulations it has EVER run (227 -> 9.4 simulations/s), and a
// whole match of full games never finishes.  The search itself (what tier 3 compares) is untouched; the CPU-baseline build
// (libref_mcts.so, one fresh player per timed decision) keeps the reference's own allocator.
#pragma once
#include <cstdint>
#include <cstddef>
#include <new>
#include <unordered_set>

template <int CHUNK_SIZE, int NUM_CHUNKS_PER_BLOCK>
struct ObjectPoolMultisize
{
	static const int NumChunksPerBlock = NUM_CHUNKS_PER_BLOCK;
	static const int ChunkSize = CHUNK_SIZE;
	~ObjectPoolMultisize() { free_all(); }
	uint8_t* alloc(size_t num_chunks)
	{
		tot_current_usage += num_chunks;
		if (max_usage < tot_current_usage) max_usage = tot_current_usage;
		uint8_t* p = static_cast<uint8_t*>(::operator new(num_chunks * CHUNK_SIZE));
		live.insert(p);
		return p;
	}
	template <typename T, typename... Args> T* alloc(size_t num_chunks, Args... args) { return new (alloc(num_chunks)) T(args...); }
	void free(uint8_t* ptr, size_t num_chunks) { tot_current_usage -= num_chunks; live.erase(ptr); ::operator delete(ptr); }
	template <typename T> void free(T* ptr, size_t num_chunks) { ptr->~T(); free(reinterpret_cast<uint8_t*>(ptr), num_chunks); }
	void free_all() { for (uint8_t* p : live) ::operator delete(p); live.clear(); tot_current_usage = 0; }
	void reset_stats() { max_usage = 0; }
	size_t get_max_usage() { return max_usage; }
	size_t get_current_usage() { return tot_current_usage; }
protected:
	std::unordered_set<uint8_t*> live;
	size_t max_usage = 0, tot_current_usage = 0;
};
