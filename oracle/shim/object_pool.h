// oracle/shim/object_pool.h -- TEST INFRASTRUCTURE ONLY.
// The reference's common/utils/object_pool.h is MSVC-only (object_pool.h:76 lacks the
// `template` disambiguator, :112 leaves a parameter pack unexpanded).  The rules code only
// needs ObjectPoolBlocked<T,N>::alloc()/free(); this stand-in is shadowing it by include order.
#pragma once
#include <cstdint>
#include <vector>
#include <new>
template <typename T, int OBJECTS_IN_BLOCK>
struct ObjectPoolBlocked
{
	std::vector<T*> free_list;
	std::vector<uint8_t*> blocks;
	~ObjectPoolBlocked() { for (auto* b : blocks) delete[] b; }
	T* alloc()
	{
		if (free_list.empty()) {
			auto* blk = new uint8_t[sizeof(T) * OBJECTS_IN_BLOCK];
			blocks.push_back(blk);
			for (int i = 0; i < OBJECTS_IN_BLOCK; ++i) free_list.push_back(reinterpret_cast<T*>(blk + i * sizeof(T)));
		}
		T* p = free_list.back(); free_list.pop_back();
		return new (p) T();
	}
	void free(T* p) { free_list.push_back(p); }
};
