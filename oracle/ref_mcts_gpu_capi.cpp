// oracle/ref_mcts_gpu_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// INTEGRATION.md section 2 as a binary: the reference's OWN MC::Player (c++/MCTSPlayer/MCTSPlayer.{h,cpp}) with
// oracle/mcts_gpu_stub.patch applied -- its simulation loop cut at the first new state, the playouts from those leaves
// played by libgameai_gpu.so through the C ABI (gai_create / gai_loa_rollout), back-up and expansion unchanged -- linked
// with the reference's own LOA rules.  tests/test_gpu_integration_stub.py plays moves with it on a GPU.
#include REF_LOA_CPP
#include REF_MCTS_CPP
#include REF_FACTORY_CPP
#include <chrono>
#include <random>

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

// The patched player as Blacks from the opening against a uniformly random mover (fixed seed), `n_moves` move decisions of
// `sims` simulations each.  out_moves: from/to pairs of the player's moves; *gpu_playouts: playouts the GPU ran for it.
// Returns the number of moves played (the game may end earlier), or -1 when the GPU library refuses (no device).
int ref_mcts_gpu_play(int n_moves, int sims, unsigned seed, uint8_t* out_moves, unsigned long long* gpu_playouts, double* seconds, char* err, int errlen)
{
	try {
		LinesOfAction::LinesOfActionGameRules rules; rules.m_RefCnt = 1000;
		PlayerConfig_t pc;
		pc.put("name", "mcts"); pc.put("type", "mcts"); pc.put("number_of_players", 2);
		pc.put("random_seed", seed); pc.put("explore_exploit_ratio", 2.0); pc.put("playout_depth", 50000);
		pc.put("expand_size", 3); pc.put("best_move_value_eps", 0.05); pc.put("move_sim_limit", sims);
		IGamePlayer* p = MC::createMCPlayer(1, pc);
		p->setGameRules(&rules);                            // gai_create happens here
		GameState* gs = rules.CreateRandomInitialState(nullptr);
		p->startNewGame(gs);
		std::mt19937 rng(seed * 7919u + 1);
		const auto t0 = std::chrono::steady_clock::now();
		int played = 0;
		while (played < n_moves && !rules.IsTerminal(gs)) {
			MoveList* ml;
			if (rules.GetCurrentPlayer(gs) == 1) {
				ml = p->selectMove(gs);
				out_moves[2 * played] = ml->move[0].from; out_moves[2 * played + 1] = ml->move[0].to;
				++played;
			} else {
				MoveList* all = rules.GetPlayerLegalMoves(gs, 0);
				ml = rules.SelectMoveFromList(all, (int)(rng() % (unsigned)rules.GetNumMoves(all)));
				rules.ReleaseMoveList(all);
			}
			GameState* ns = rules.ApplyMove(gs, &ml->move[0], rules.GetCurrentPlayer(gs));
			rules.ReleaseMoveList(ml);
			rules.ReleaseGameState(gs);
			gs = ns;
		}
		*seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		*gpu_playouts = static_cast<MC::Player*>(p)->m_gpuPlayouts;
		p->endGame(0, GameResult::Draw);
		p->release();
		rules.ReleaseGameState(gs);
		return played;
	} catch (const char* msg) {
		snprintf(err, (size_t)errlen, "%s", msg);
		return -1;
	}
}

} // extern "C"
