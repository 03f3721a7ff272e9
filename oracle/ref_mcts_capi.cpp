// oracle/ref_mcts_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN MCTS player: c++/MCTSPlayer/MCTSPlayer.cpp (UTF-16 in the reference tree, transcoded by the
// Makefile into the git-ignored oracle/_ref/ with the one MSVC-only cast at :92 rewritten) + MCplayer_factory.cpp,
// linked with the reference's own LOA rules, behind a flat C API.  Used for the tier-3 comparison (move choice on
// tactical positions) and as the "reference MCTSPlayer" CPU baseline (simulations/s through selection_playOut).
// Not compiled: MCTSPlayer_diagnostics.cpp and MCRLplayer.cpp (boost::xpressive) -- their entry points are stubbed.
#include REF_LOA_CPP
#include REF_MCTS_CPP
#include REF_FACTORY_CPP
#include <chrono>

// ---- stubs for the parts that are not built ---------------------------------------------------------
namespace Trace {
struct NullTrace : ITrace { void trace(const wchar_t*) override {} void trace(const wchar_t*, ParamListGen_t) override {} void release() override { delete this; } };
ITrace* createInstance(const string&, const string&) { return new NullTrace(); }
}
namespace MC {
void Player::_dumpMoveTree() {}
void Player::_dumpGameTree() {}
void Player::dumpTree(const string&, StateNode*, std::set<StateNode*>) {}
bool Player::checkTree(StateNode*, unsigned short, bool) { return true; }
}
namespace MCRL { IGamePlayer* createMCRLPlayer(int, const PlayerConfig_t&) { return nullptr; } }

extern "C" {

// One move decision of the reference MC::Player on the LOA position (w, b, cp) with a fixed simulation budget.
// out_move[2] = chosen from/to; visits/values: root children in move-list order; returns the number of root moves.
static int select_impl(uint64_t w, uint64_t b, int cp, int sims, double time_limit, unsigned seed, double C, int playout_depth, int expand_size, double eps,
                       uint8_t* out_move, uint32_t* visits, float* values, double* seconds, double* mean_path, long* runs);
int ref_mcts_select(uint64_t w, uint64_t b, int cp, int sims, unsigned seed, double C, int playout_depth, int expand_size, double eps,
                    uint8_t* out_move, uint32_t* visits, float* values, double* seconds, double* mean_path)
{
	long runs = 0;
	return select_impl(w, b, cp, sims, 0.0, seed, C, playout_depth, expand_size, eps, out_move, visits, values, seconds, mean_path, &runs);
}
// the same decision under the reference's TimeMoveLimit (move_time_limit, MCplayer_factory.cpp:23-35): *runs = simulations it managed
int ref_mcts_select_timed(uint64_t w, uint64_t b, int cp, double time_limit, unsigned seed, double C, int playout_depth, int expand_size, double eps,
                          uint8_t* out_move, uint32_t* visits, float* values, double* seconds, double* mean_path, long* runs)
{
	return select_impl(w, b, cp, 0, time_limit, seed, C, playout_depth, expand_size, eps, out_move, visits, values, seconds, mean_path, runs);
}
static int select_impl(uint64_t w, uint64_t b, int cp, int sims, double time_limit, unsigned seed, double C, int playout_depth, int expand_size, double eps,
                       uint8_t* out_move, uint32_t* visits, float* values, double* seconds, double* mean_path, long* runs)
{
	LinesOfAction::LinesOfActionGameRules rules; rules.m_RefCnt = 1000;
	PlayerConfig_t pc;
	pc.put("name", "mcts"); pc.put("type", "mcts"); pc.put("number_of_players", 2);
	pc.put("random_seed", seed); pc.put("explore_exploit_ratio", C); pc.put("playout_depth", playout_depth);
	pc.put("expand_size", expand_size); pc.put("best_move_value_eps", eps);
	if (time_limit > 0) pc.put("move_time_limit", time_limit); else pc.put("move_sim_limit", sims);
	IGamePlayer* p = MC::createMCPlayer(cp, pc);
	p->setGameRules(&rules);
	GameState gs(w, b, cp);
	p->startNewGame(&gs);
	const auto t0 = std::chrono::steady_clock::now();
	MoveList* ml = p->selectMove(&gs);
	*seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	out_move[0] = ml->move[0].from; out_move[1] = ml->move[0].to;
	auto* mp = static_cast<MC::Player*>(p);
	MC::StateNode* root = mp->m_root;
	const int nm = root->numMoves;
	for (int i = 0; i < nm; ++i) { visits[i] = root->moves[i].numVisited; values[2 * i] = root->moves[i].value[0]; values[2 * i + 1] = root->moves[i].value[1]; }
	double tot = 0, cnt = 0;
	for (auto& kv : mp->m_path_size.values) { tot += (double)kv.first * kv.second; cnt += kv.second; }
	*mean_path = cnt ? tot / cnt : 0;
	*runs = (long)cnt;                                 // one path_size entry per simulation (MCTSPlayer.cpp:468)
	rules.ReleaseMoveList(ml);
	p->endGame(0, GameResult::Draw);
	p->release();
	return nm;
}

} // extern "C"
