// oracle/shim/boost/bimap.hpp -- TEST INFRASTRUCTURE ONLY.
// MC::Player keeps a string <-> StateNode* bimap (MCTSPlayer.h:143) and uses: left.find / left.insert / left.end,
// right.find / right.erase / right.end, size, clear (MCTSPlayer.cpp:132-134,641-659,690-694).
#pragma once
#include <map>
#include <utility>
namespace boost {
namespace bimaps { template <class T> struct set_of { using type = T; }; }
template <class LS, class RS> struct bimap {
	using L = typename LS::type; using R = typename RS::type;
	struct left_view {
		bimap* bm; std::map<L, R> m;
		using iterator = typename std::map<L, R>::iterator;
		iterator find(const L& k) { return m.find(k); }
		iterator end() { return m.end(); }
		std::pair<iterator, bool> insert(const std::pair<L, R>& p)
		{
			auto r = m.insert(p);
			if (r.second) bm->right.m.insert({ p.second, p.first });
			return r;
		}
	} left;
	struct right_view {
		bimap* bm; std::map<R, L> m;
		using iterator = typename std::map<R, L>::iterator;
		iterator find(const R& k) { return m.find(k); }
		iterator end() { return m.end(); }
		void erase(iterator it) { bm->left.m.erase(it->second); m.erase(it); }
	} right;
	bimap() { left.bm = this; right.bm = this; }
	size_t size() const { return left.m.size(); }
	void clear() { left.m.clear(); right.m.clear(); }
};
}
