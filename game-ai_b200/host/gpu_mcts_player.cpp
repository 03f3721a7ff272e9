// game-ai_b200/host/gpu_mcts_player.cpp -- GPU-backed MCTS player plugin, exports createPlayer.
//
// Drop-in variant of the reference's MC::Player (c++/MCTSPlayer/MCTSPlayer.{h,cpp}, factory MCplayer_factory.cpp:46-82):
// same IGamePlayer vtable, same config keys, same tree policy --
//   edge value  = value[player] * probability / (1 + visits)                      (MoveNode::getWeight, MCTSPlayer.h:55-58)
//   UCB         = value + C * sqrt(ln(N + 1) / (1 + visits)), ties broken uniformly (Player::selectMove, MCTSPlayer.cpp:564-599)
//   final move  = best VALUE, uniformly among those within best_move_value_eps       (selectBestMove, :601-632)
//   states reached by different move orders share one node                          (makeTreeNode, :634-662)
// but the simulation phase (the part of selection_playOut, :387-470, below the tree) is a batch of leaves sent to
// libgameai_gpu.so (gameai_gpu.h gai_loa_rollout): selection/expansion/back-propagation stay here on the host.
// Differences that follow from batching, all deliberate: leaves are collected with a virtual loss (the edge is charged
// its playouts before their values arrive), every selected leaf becomes a permanent node, visit counters are 32-bit
// (the reference's unsigned char wraps at 255, MCTSPlayer.h:51), and draws/capped playouts back up 0.5 like the
// reference's Score/eval (LinesOfActionZasady.cpp:382-397,420-432).
//
// Extra config keys: gpu_batch (leaves per launch), playouts_per_leaf, devices ("0" or "0,1,2,3": root parallelism,
// one independent tree per GPU merged at the root), philox_seed.  Multi-process root parallelism (one rank per GPU):
// WORLD_SIZE / RANK / GAI_NCCL_ID_FILE in the environment -> gai_allreduce_root over NCCL once per move decision.
#include "gameai/GamePlayer.h"
#include "gameai/GameRules.h"
#include "gameai_gpu.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

using Clock = std::chrono::steady_clock;

struct Edge { int32_t child = -1; uint32_t visits = 0; float value[2] = { 0.f, 0.f }; };
struct Node {
	GameState* state = nullptr; MoveList* moves = nullptr;
	int first_edge = 0, n_edges = -1;        // n_edges < 0: moves not generated yet (most leaves are never descended into)
	uint32_t visits = 0, visit_id = 0;
	uint8_t cp = 0; bool terminal = false;
	int score[2] = { 0, 0 };
};
struct Key { uint64_t w, b, cp; bool operator==(const Key& o) const { return w == o.w && b == o.b && cp == o.cp; } };
struct KeyHash { size_t operator()(const Key& k) const { uint64_t h = k.w * 0x9E3779B97F4A7C15ull ^ (k.b + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull ^ k.cp; h ^= h >> 29; return (size_t)h; } };

// Transposition table (makeTreeNode, MCTSPlayer.cpp:634-662, keyed there by the ToString text): open addressing, one
// 32-byte entry per position, no allocation per insert (std::unordered_map cost two cache misses and a malloc per node).
struct FlatTable {
	struct Entry { Key k; int32_t idx; uint32_t used; };
	std::vector<Entry> slots; size_t mask = 0, count = 0;
	void init(size_t cap) { size_t c = 1; while (c < cap) c <<= 1; slots.assign(c, Entry{ Key{ 0, 0, 0 }, -1, 0 }); mask = c - 1; count = 0; }
	void clear() { if (!slots.empty()) init(slots.size()); }
	int find(const Key& k) const
	{
		if (slots.empty()) return -1;
		for (size_t i = KeyHash()(k) & mask;; i = (i + 1) & mask) {
			const Entry& e = slots[i];
			if (!e.used) return -1;
			if (e.k == k) return e.idx;
		}
	}
	void emplace(const Key& k, int idx)
	{
		if (slots.empty()) init(1u << 19);
		if (2 * (count + 1) > slots.size()) {                                // keep the load factor under 1/2
			std::vector<Entry> old; old.swap(slots);
			init(old.size() * 2);
			for (const Entry& e : old) if (e.used) emplace(e.k, e.idx);
		}
		for (size_t i = KeyHash()(k) & mask;; i = (i + 1) & mask) {
			Entry& e = slots[i];
			if (!e.used) { e.k = k; e.idx = idx; e.used = 1; ++count; return; }
			if (e.k == k) { e.idx = idx; return; }
		}
	}
};

struct Settings {
	int player = 0, num_players = 2;
	float C = 1.0f, eps = 0.005f;
	int max_plies = 20;
	int batch = 1024, ppl = 32;
	double time_limit = 1.0; long sim_limit = -1;
	uint64_t philox_key = 0x10A5EED;
	unsigned seed = 1;
	std::string name;
};

// One independent search tree bound to one GPU context (root parallelism = several of these).
struct Tree {
	IGameRules* rules = nullptr;
	gai_ctx* ctx = nullptr;
	Settings cfg;
	std::vector<Node> nodes;
	std::vector<Edge> edges;
	FlatTable table;
	std::mt19937 rng;
	uint32_t visit_id = 0;
	uint64_t next_playout_id = 0, key = 0;
	// per-search statistics
	uint64_t playouts = 0, plies = 0, batches = 0, launches = 0;
	double gpu_ms = 0;
	std::string error;

	Key key_of(const GameState* s) const
	{
		const uint32_t* h = rules->GetStateHash(s);          // LOA: the 24-byte {white, black, current_player}
		Key k; memcpy(&k, h, 24); k.cp &= 1; return k;
	}
	int make_node(GameState* s)                               // takes ownership of s
	{
		const Key k = key_of(s);
		const int known = table.find(k);
		if (known >= 0) { rules->ReleaseGameState(s); return known; }
		if (nodes.capacity() == 0) { nodes.reserve(1u << 18); edges.reserve(1u << 21); }      // no regrowth mid-search
		Node n; n.state = s; n.cp = (uint8_t)rules->GetCurrentPlayer(s);
		n.terminal = rules->IsTerminal(s);
		if (n.terminal) { rules->Score(s, n.score); n.n_edges = 0; }
		nodes.push_back(n);
		const int idx = (int)nodes.size() - 1;
		table.emplace(k, idx);
		return idx;
	}
	// Move generation is deferred to the first descent INTO a node: every selection ends by creating one new leaf, and
	// most leaves are never selected again -- GetPlayerLegalMoves + the edge array were 45 % of the host time per selection.
	void expand(Node& n)
	{
		if (n.n_edges >= 0) return;
		n.moves = rules->GetPlayerLegalMoves(n.state, n.cp);
		n.n_edges = rules->GetNumMoves(n.moves);
		n.first_edge = (int)edges.size();
		edges.resize(edges.size() + (size_t)n.n_edges);
	}
	void clear()
	{
		for (auto& n : nodes) { if (n.moves) rules->ReleaseMoveList(n.moves); rules->ReleaseGameState(n.state); }
		nodes.clear(); edges.clear(); table.clear();
	}
	// Tree reuse with garbage collection (reference: findRootNode + freeSubTree, MCTSPlayer.cpp:219-251,330-354):
	// keep what is reachable from the new root, free the rest, compact the arenas.  Returns the new root index.
	int collect(int root)
	{
		std::vector<int> remap(nodes.size(), -1), order;
		order.reserve(nodes.size() / 4 + 16);
		remap[root] = 0; order.push_back(root);
		for (size_t h = 0; h < order.size(); ++h) {                       // breadth first over the DAG
			const Node& n = nodes[order[h]];
			for (int i = 0; i < n.n_edges; ++i) {
				const int c = edges[n.first_edge + i].child;
				if (c >= 0 && remap[c] < 0) { remap[c] = (int)order.size(); order.push_back(c); }
			}
		}
		std::vector<Node> nn; nn.reserve(std::max(order.size() * 2, (size_t)1 << 18));
		std::vector<Edge> ne; ne.reserve(std::max(edges.size() / 2 + 64, (size_t)1 << 21));
		for (int old : order) {
			Node n = nodes[old];
			const int fe = (int)ne.size();
			for (int i = 0; i < n.n_edges; ++i) { Edge e = edges[n.first_edge + i]; if (e.child >= 0) e.child = remap[e.child]; ne.push_back(e); }
			n.first_edge = fe; n.visit_id = 0;
			nn.push_back(n);
		}
		for (size_t i = 0; i < nodes.size(); ++i) if (remap[i] < 0) { if (nodes[i].moves) rules->ReleaseMoveList(nodes[i].moves); rules->ReleaseGameState(nodes[i].state); }
		nodes.swap(nn); edges.swap(ne);
		table.clear(); visit_id = 0;
		for (size_t i = 0; i < nodes.size(); ++i) table.emplace(key_of(nodes[i].state), (int)i);
		return 0;
	}

	// UCB choice among the moves that do not lead back into the current path (MCTSPlayer.cpp:564-599).  The values live in
	// the node's contiguous edge array; whether a child lies on the current path is checked only for the edge that WON
	// (then that edge is excluded and the choice repeated): reading every child's visit stamp cost one cache miss per
	// edge per level -- 5 of the 7 microseconds a selection took on a 200 k-node tree.
	int choose_edge(const Node& n)
	{
		const float lnN = (float)std::log((double)n.visits + 1.0);
		const int ne = n.n_edges < 128 ? n.n_edges : 128;                 // <= 96 moves (LinesOfActionZasady.cpp:226)
		float val[128];
		const Edge* ed = edges.data() + n.first_edge;
		const int cp = n.cp;
		for (int i = 0; i < ne; ++i) {                                    // no branches: the compiler vectorises sqrt and divide (8 lanes)
			const float oo = 1.0f / (1.0f + (float)ed[i].visits);
			val[i] = ed[i].value[cp] * oo + cfg.C * std::sqrt(lnN * oo);
		}
		for (;;) {
			float best = -1e30f; int ties[128]; int nt = 0;
			for (int i = 0; i < ne; ++i) {
				const float v = val[i];
				if (v > best) { best = v; nt = 0; ties[nt++] = i; }
				else if (v == best) ties[nt++] = i;
			}
			if (nt == 0 || best == -1e30f) return -1;
			const int pick = ties[nt == 1 ? 0 : std::uniform_int_distribution<int>(0, nt - 1)(rng)];
			const int child = ed[pick].child;
			if (child < 0 || nodes[child].visit_id != visit_id) return pick;
			val[pick] = -1e30f;                                           // would close a loop: choose again without it
		}
	}

	struct Pending { std::vector<int> path; int leaf; };       // path = edge indices from the root

	void backprop(const std::vector<int>& path, float vw, float vb)
	{
		for (int ei : path) { edges[ei].value[0] += vw; edges[ei].value[1] += vb; }
	}
	// one selection; charges `ppl` virtual visits along the path.  Returns false if the simulation was resolved on the
	// host (terminal node or dead end) and needs no playout.
	bool select_leaf(int root, Pending& out)
	{
		++visit_id;
		out.path.clear();
		int ni = root;
		nodes[ni].visits += cfg.ppl;
		for (;;) {
			Node& n = nodes[ni];
			n.visit_id = visit_id;
			if (n.terminal) { backprop(out.path, n.score[0] / 100.0f * cfg.ppl, n.score[1] / 100.0f * cfg.ppl); return false; }
			expand(n);
			const int ci = choose_edge(n);
			if (ci < 0) { backprop(out.path, 0.5f * cfg.ppl, 0.5f * cfg.ppl); return false; }      // cycle_score 50 (MCTSConfig::CycleScore)
			const int ei = n.first_edge + ci;
			edges[ei].visits += cfg.ppl;
			out.path.push_back(ei);
			if (edges[ei].child >= 0) { ni = edges[ei].child; nodes[ni].visits += cfg.ppl; continue; }
			auto [mv, prob] = rules->GetMoveFromList(n.moves, ci);
			(void)prob;
			GameState* ns = rules->ApplyMove(n.state, mv, n.cp);
			const size_t before = nodes.size();
			const int child = make_node(ns);                      // may reallocate `nodes`: n is dead after this line
			edges[ei].child = child;
			nodes[child].visits += cfg.ppl;
			if (nodes.size() == before) {                         // transposition into an existing node: keep descending
				if (nodes[child].visit_id == visit_id) { backprop(out.path, 0.5f * cfg.ppl, 0.5f * cfg.ppl); return false; }
				ni = child; continue;
			}
			if (nodes[child].terminal) { backprop(out.path, nodes[child].score[0] / 100.0f * cfg.ppl, nodes[child].score[1] / 100.0f * cfg.ppl); return false; }
			out.leaf = child;
			return true;
		}
	}

	// Run simulations from `root` until the limit.  Two batches are pipelined: while batch A plays out on the GPU
	// (gai_loa_rollout_async), the host selects batch B -- A's leaves already carry their virtual loss, so B explores
	// elsewhere -- then A's results are backed up and B is launched.
	struct Batch {
		std::vector<Pending> pend; std::vector<gai_loa_state> leaves; std::vector<gai_loa_result> res; int nl = 0;
		explicit Batch(int b) : pend((size_t)b), leaves((size_t)b), res((size_t)b) {}
	};
	long fill_batch(int root, Batch& b)
	{
		b.nl = 0;
		for (int k = 0; k < cfg.batch; ++k)
			if (select_leaf(root, b.pend[b.nl])) { memcpy(&b.leaves[b.nl], rules->GetStateHash(nodes[b.pend[b.nl].leaf].state), 24); ++b.nl; }
		return (long)cfg.batch * cfg.ppl;
	}
	bool launch(Batch& b)
	{
		if (b.nl == 0) return true;
		const int rc = gai_loa_rollout_async(ctx, b.leaves.data(), (uint32_t)b.nl, (uint32_t)cfg.ppl, (uint32_t)cfg.max_plies, key, next_playout_id, b.res.data());
		if (rc != GAI_OK) { error = gai_last_error(ctx); return false; }
		next_playout_id += (uint64_t)b.nl * cfg.ppl;
		return true;
	}
	bool finish(Batch& b)
	{
		if (b.nl == 0) return true;
		if (gai_sync(ctx) != GAI_OK) { error = gai_last_error(ctx); return false; }
		gai_rollout_stats st;
		if (gai_last_rollout_stats(ctx, &st) == GAI_OK) { playouts += st.playouts; plies += st.plies; gpu_ms += st.total_ms; }
		++batches;
		for (int i = 0; i < b.nl; ++i) {
			const float half = 0.5f * (b.res[i].draws + b.res[i].capped);
			backprop(b.pend[i].path, b.res[i].white_wins + half, b.res[i].black_wins + half);
		}
		b.nl = 0;
		return true;
	}
	void search(int root, Clock::time_point deadline, long sim_limit)
	{
		Batch A(cfg.batch), B(cfg.batch);
		Batch* fly = &A; Batch* next = &B;
		long sims = fill_batch(root, *fly);
		if (!launch(*fly)) return;
		while (sim_limit > 0 ? sims < sim_limit : Clock::now() < deadline) {
			sims += fill_batch(root, *next);          // overlaps the GPU run of *fly
			if (!finish(*fly)) return;
			if (!launch(*next)) return;
			std::swap(fly, next);
		}
		finish(*fly);
	}
};

struct GpuMctsPlayer : IGamePlayer
{
	Settings cfg;
	IGameRules* rules = nullptr;
	std::vector<Tree> trees;                 // one per device
	std::vector<int> devices;
	std::mt19937 tie_rng;                    // final-move tie break: identical on every rank
	Histogram<long> runs_per_move, branching;
	Average playouts_per_sec, plies_per_playout, gpu_busy;
	double total_batches = 0;
	int world = 1, rank = 0; bool comm_ready = false;
	int move_nbr = 0;
	std::string last_root;                   // "visits:value_white:value_black,..." of the last move decision (merged over trees/ranks)
	double gc_runs = 0, gc_freed = 0, reused_visits = 0;

	GpuMctsPlayer(const Settings& s, const std::vector<int>& devs) : cfg(s), devices(devs), tie_rng(s.seed) { runs_per_move.Bucket(1000); }
	~GpuMctsPlayer() override { for (auto& t : trees) { if (rules) t.clear(); if (t.ctx) gai_destroy(t.ctx); } }

	void release() override { delete this; }
	std::string getName() override { return cfg.name; }
	void resetStats() override {}
	void startNewGame(GameState*) override { move_nbr = 0; }
	void endGame(int, GameResult) override { for (auto& t : trees) t.clear(); }

	void setGameRules(IGameRules* gr) override
	{
		rules = gr;
		if (gr->GetStateHashSize() != 6) throw "gpumctsplayer: the GPU rollout path is built for the Lines of Action rules (24-byte state)";
		if (const char* ws = getenv("WORLD_SIZE")) world = atoi(ws);
		if (const char* rk = getenv("RANK")) rank = atoi(rk);
		trees.resize(devices.size());
		for (size_t i = 0; i < devices.size(); ++i) {
			Tree& t = trees[i];
			t.rules = gr; t.cfg = cfg; t.rng.seed(cfg.seed + 7919u * (unsigned)(i + (size_t)rank * devices.size()));
			t.key = cfg.philox_key ^ (0x9E3779B97F4A7C15ull * (uint64_t)(i + (size_t)rank * devices.size()));   // own Philox stream per tree
			if (gai_create(devices[i], &t.ctx) != GAI_OK) {
				static std::string msg; msg = std::string("gpumctsplayer: ") + gai_last_error(nullptr);
				throw msg.c_str();                             // no CPU fallback: the player cannot play without its GPU
			}
		}
		if (world > 1) init_comm();
	}
	// rank 0 writes the NCCL unique id to GAI_NCCL_ID_FILE, the other ranks poll for it
	void init_comm()
	{
		const char* path = getenv("GAI_NCCL_ID_FILE");
		if (!path) throw "gpumctsplayer: WORLD_SIZE > 1 needs GAI_NCCL_ID_FILE";
		uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES];
		if (rank == 0) {
			if (gai_comm_unique_id(id) != GAI_OK) throw "gpumctsplayer: ncclGetUniqueId failed";
			std::string tmp = std::string(path) + ".tmp";
			{ std::ofstream f(tmp, std::ios::binary); f.write((const char*)id, sizeof id); }
			rename(tmp.c_str(), path);
		} else {
			for (int tries = 0;; ++tries) {
				std::ifstream f(path, std::ios::binary);
				if (f && f.read((char*)id, sizeof id)) break;
				if (tries > 600) throw "gpumctsplayer: timed out waiting for the NCCL id file";
				std::this_thread::sleep_for(std::chrono::milliseconds(100));
			}
		}
		if (gai_comm_init(trees[0].ctx, rank, world, id) != GAI_OK) throw "gpumctsplayer: gai_comm_init failed";
		comm_ready = true;
	}

	MoveList* selectMove(GameState* pks) override
	{
		if (rules->GetCurrentPlayer(pks) != cfg.player) return rules->GetPlayerLegalMoves(pks, cfg.player);     // the shared noop list
		++move_nbr;
		const auto t0 = Clock::now();
		const auto deadline = t0 + std::chrono::microseconds((long)(cfg.time_limit * 1e6));
		std::vector<int> roots(trees.size());
		for (size_t i = 0; i < trees.size(); ++i) {
			Tree& t = trees[i];
			roots[i] = t.make_node(rules->CopyGameState(pks));                        // reuses the subtree when the state is known
			if (t.nodes.size() > 200000) { const size_t before = t.nodes.size(); roots[i] = t.collect(roots[i]); gc_freed += before - t.nodes.size(); ++gc_runs; }
			t.playouts = t.plies = t.batches = 0; t.gpu_ms = 0;
			reused_visits += t.nodes[roots[i]].visits;                               // > 0 when the position was found in the old tree
		}
		// rules pools are not thread safe (GameController.cpp:167-180): extra trees get their own rules instance? No --
		// the host rules plugin here allocates per call, so one instance is shared read-mostly; searches run one thread per tree.
		if (trees.size() == 1) trees[0].search(roots[0], deadline, cfg.sim_limit);
		else {
			std::vector<std::thread> th;
			for (size_t i = 0; i < trees.size(); ++i) th.emplace_back([&, i] { trees[i].search(roots[i], deadline, cfg.sim_limit > 0 ? cfg.sim_limit / (long)trees.size() : -1); });
			for (auto& x : th) x.join();
		}
		for (auto& t : trees) if (!t.error.empty()) { static std::string msg; msg = "gpumctsplayer: " + t.error; throw msg.c_str(); }

		// merge the root statistics of all trees (children are index aligned: the move list order is deterministic)
		const Node& r0 = trees[0].nodes[roots[0]];
		const int nm = r0.n_edges;
		std::vector<uint32_t> visits((size_t)nm, 0); std::vector<float> values((size_t)nm * 2, 0.f);
		uint64_t playouts = 0, plies = 0, batches = 0; double gpu_ms = 0;
		for (size_t i = 0; i < trees.size(); ++i) {
			const Node& r = trees[i].nodes[roots[i]];
			for (int m = 0; m < nm; ++m) {
				const Edge& e = trees[i].edges[r.first_edge + m];
				visits[m] += e.visits; values[2 * m] += e.value[0]; values[2 * m + 1] += e.value[1];
			}
			playouts += trees[i].playouts; plies += trees[i].plies; batches += trees[i].batches; gpu_ms = std::max(gpu_ms, trees[i].gpu_ms);
		}
		if (comm_ready && nm > 0 && gai_allreduce_root(trees[0].ctx, visits.data(), values.data(), (uint32_t)nm) != GAI_OK) throw "gpumctsplayer: gai_allreduce_root failed";

		// final move: best value, uniformly among those within eps (MCTSPlayer.cpp:601-618)
		std::vector<int> best; float bestv = 0;
		for (int m = 0; m < nm; ++m) {
			const float v = values[2 * m + cfg.player] / (1.0f + visits[m]);
			if (best.empty()) { best.push_back(m); bestv = v; continue; }
			const float diff = v - bestv;
			if (std::fabs(diff) < cfg.eps) best.push_back(m);
			else if (diff > 0) { best.clear(); best.push_back(m); bestv = v; }
		}
		const int pick = best.empty() ? 0 : best[best.size() == 1 ? 0 : std::uniform_int_distribution<size_t>(0, best.size() - 1)(tie_rng)];
		{
			std::ostringstream rs;
			for (int m = 0; m < nm; ++m) rs << (m ? "," : "") << visits[m] << ":" << values[2 * m] << ":" << values[2 * m + 1];
			last_root = rs.str();
		}
		const double secs = std::chrono::duration<double>(Clock::now() - t0).count();
		runs_per_move.insert((long)playouts);
		branching.insert(nm);
		if (secs > 0) playouts_per_sec.insert(playouts / secs);
		if (playouts) plies_per_playout.insert((double)plies / playouts);
		if (secs > 0) gpu_busy.insert(gpu_ms * 1e-3 / secs);
		total_batches += batches;
		return rules->SelectMoveFromList(r0.moves, pick);
	}

	NamedMetrics_t getGameStats() override
	{
		NamedMetrics_t nm;
		nm["runs_per_move"] = runs_per_move;                   // same name as the reference (MCTSPlayer.cpp:152)
		nm["branching_factor"] = branching;
		nm["playouts_per_sec"] = playouts_per_sec;
		nm["plies_per_playout"] = plies_per_playout;
		nm["gpu_busy_fraction"] = gpu_busy;
		nm["gpu_batches"] = total_batches;
		nm["last_root_stats"] = last_root;
		nm["tree_gc_runs"] = gc_runs; nm["tree_gc_freed_nodes"] = gc_freed; nm["reused_root_visits"] = reused_visits;
		return nm;
	}
};

} // namespace

extern "C" IGamePlayer* createPlayer(int player_number, const PlayerConfig_t* pc)
{
	Settings s;
	s.player = player_number;
	s.seed = pc->get_optional<unsigned>("random_seed").get_value_or((unsigned)Clock::now().time_since_epoch().count());
	s.num_players = pc->get_optional<int>("number_of_players").get_value_or(2);
	s.max_plies = pc->get_optional<int>("playout_depth").get_value_or(20);
	s.C = pc->get_optional<float>("explore_exploit_ratio").get_value_or(1.0f);
	s.eps = pc->get_optional<float>("best_move_value_eps").get_value_or(0.005f);
	s.batch = std::max(1, pc->get_optional<int>("gpu_batch").get_value_or(1024));
	s.ppl = std::max(1, pc->get_optional<int>("playouts_per_leaf").get_value_or(32));
	s.philox_key = pc->get_optional<uint64_t>("philox_seed").get_value_or((uint64_t)s.seed * 0x9E3779B97F4A7C15ull + 1);
	auto sl = pc->get_optional<long>("move_sim_limit");
	if (sl) s.sim_limit = sl.get(); else s.time_limit = pc->get_optional<double>("move_time_limit").get_value_or(1.0);
	s.name = pc->get_optional<std::string>("fullname").get_value_or(pc->get_optional<std::string>("name").get_value_or("gpumcts"));
	std::vector<int> devs;
	{
		std::string d = pc->get_optional<std::string>("devices").get_value_or("");
		if (d.empty()) { const char* lr = getenv("LOCAL_RANK"); devs.push_back(lr ? atoi(lr) : 0); }
		else { std::replace(d.begin(), d.end(), ',', ' '); std::istringstream ss(d); int x; while (ss >> x) devs.push_back(x); }
	}
	return new GpuMctsPlayer(s, devs);
}
