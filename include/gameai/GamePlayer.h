// include/gameai/GamePlayer.h -- the player plugin interface, same shape as the reference's
// c++/common/api/GamePlayer.h:12-30.  Call order used by the controller (GameController.cpp:54-61,104,199-204):
// create -> setGameRules -> per game startNewGame -> every round selectMove for EVERY player (the one not to move
// must return the rules' noop list) -> endGame -> getGameStats -> release.
#pragma once
#include <string>
#include "Config.h"
#include "Metrics.h"

struct MoveList;
struct GameState;
struct IGameRules;
using PlayerConfig_t = gai::Config;

enum class GameResult { Win = 0, Draw, Loose, AbortedByRoundLimit = 1, AbortedByStateLoop = 2 };   // GamePlayer.h:11
struct IGamePlayer
{
	virtual void            startNewGame(GameState* playerKnownState) = 0;
	virtual void            endGame(int score, GameResult result) = 0;
	virtual void            setGameRules(IGameRules*) = 0;
	virtual MoveList*       selectMove(GameState* playerKnownState) = 0;
	virtual NamedMetrics_t  getGameStats() = 0;
	virtual std::string     getName() = 0;
	virtual void            resetStats() = 0;
	virtual void            release() = 0;

protected:
	virtual ~IGamePlayer() {}
};
// Exported by a player plugin as  extern "C" IGamePlayer* createPlayer(int player_number, const PlayerConfig_t*)
// (reference: BOOST_DLL_ALIAS(MC::createMCPlayer, createPlayer), MCplayer_factory.cpp:79-82).
using CreatePlayer_t = IGamePlayer* (*)(int player_number, const PlayerConfig_t* cfg);
