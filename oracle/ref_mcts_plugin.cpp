// oracle/ref_mcts_plugin.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN MCTS player (c++/MCTSPlayer/MCTSPlayer.cpp + MCplayer_factory.cpp, transcoded verbatim into the
// git-ignored oracle/_ref/ by the Makefile) as a player PLUGIN for this repo's GameLauncher: exports createPlayer, so that
// tier-3 matches can seat the reference player against the GPU-backed one (and against itself, as the control) through the
// same controller loop.  The reference player talks to the game only through IGameRules* (GameRules.h:14-50), whose vtable
// include/gameai/GameRules.h reproduces slot for slot, so it runs on this repo's rules plugin object unchanged.  What differs
// between the two header sets is the configuration type (boost::property_tree::ptree there, gai::Config here) and the metrics
// value type; this file converts both.  Two things are NOT the reference's in this build, both so that whole matches can
// finish, neither touching the search policy: the pool allocator (shim_plugin/object_pool_multisize.h explains why) and
// one guard against an empty candidate set in Player::selectMove (oracle/Makefile).  Nothing in the product loads this library.
#include <map>
#include <sstream>
#include <string>
#include <variant>
#include <optional>
#include <stdexcept>
#include <functional>
#include <cstdint>
#include <vector>
#include <tuple>
#include "gameai/Config.h"
namespace ours {
#include "gameai/Metrics.h"
#include "gameai/GamePlayer.h"
}
#include REF_MCTS_CPP
#include REF_FACTORY_CPP

// ---- stubs for the parts of the reference player that are not built (diagnostics need boost::xpressive) ----
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

namespace {
struct Adapter : ours::IGamePlayer
{
	::IGamePlayer* ref;
	explicit Adapter(::IGamePlayer* p) : ref(p) {}
	void startNewGame(ours::GameState* s) override { ref->startNewGame(reinterpret_cast<::GameState*>(s)); }
	void endGame(int score, ours::GameResult r) override { ref->endGame(score, static_cast<::GameResult>(static_cast<int>(r))); }
	void setGameRules(ours::IGameRules* r) override { ref->setGameRules(reinterpret_cast<::IGameRules*>(r)); }
	ours::MoveList* selectMove(ours::GameState* s) override { return reinterpret_cast<ours::MoveList*>(ref->selectMove(reinterpret_cast<::GameState*>(s))); }
	ours::NamedMetrics_t getGameStats() override
	{
		ours::NamedMetrics_t out;
		for (auto& kv : ref->getGameStats())
			out[kv.first] = std::visit([](auto&& v) -> std::string { return std::to_string(v); }, kv.second);
		return out;
	}
	std::string getName() override { return ref->getName(); }
	void resetStats() override { ref->resetStats(); }
	void release() override { ref->release(); delete this; }
};
}

extern "C" ours::IGamePlayer* createPlayer(int player_number, const gai::Config* cfg)
{
	PlayerConfig_t pc;                                     // the ptree stand-in (oracle/shim/boost/property_tree/ptree.hpp)
	for (auto& kv : cfg->attrs) if (!kv.first.empty() && kv.first[0] != '_') pc.attrs[kv.first] = kv.second;
	return new Adapter(MC::createMCPlayer(player_number, pc));
}
