// include/gameai/GameRules.h -- the rules plugin interface, same shape as the reference's
// c++/common/api/GameRules.h:14-52 (30 pure virtuals, opaque GameState / Move / MoveList defined per rules plugin),
// so that players written against the reference interface compile against this one unchanged.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <tuple>
#include <vector>

using std::string;
using std::wstring;

struct IRandomGenerator;
struct GameState;
struct Move;
struct MoveList;
using EvalFunction_t = std::function<void(const GameState*, int score[])>;

struct IGameRules
{
	virtual void        SetRandomGenerator(IRandomGenerator*) = 0;
	virtual GameState*  CreateRandomInitialState(IRandomGenerator*) = 0;
	virtual GameState*  CreateInitialStateFromHash(const uint32_t*) = 0;
	virtual GameState*  CreateStateFromString(const wstring&) = 0;
	virtual GameState*  CreateStateFromString(const string&) = 0;
	virtual GameState*  CreatePlayerKnownState(const GameState*, int playerNum) = 0;
	virtual EvalFunction_t CreateEvalFunction(const string& name) = 0;
	virtual void        UpdatePlayerKnownState(GameState* playerKnownState, const GameState* completeGameState, const std::vector<MoveList*>& playerMoves) = 0;
	virtual GameState*  CopyGameState(const GameState*) = 0;
	virtual bool        AreEqual(const GameState*, const GameState*) = 0;
	virtual void        ReleaseGameState(GameState*) = 0;
	virtual bool        IsTerminal(const GameState*) = 0;
	virtual void        Score(const GameState*, int score[]) = 0;
	virtual int         GetCurrentPlayer(const GameState*) = 0;
	virtual MoveList*   GetPlayerLegalMoves(const GameState*, int playerNum) = 0;
	virtual void        ReleaseMoveList(MoveList*) = 0;
	virtual int         GetNumMoves(const MoveList*) = 0;
	virtual std::tuple<Move*, float> GetMoveFromList(MoveList*, int idx) = 0;
	virtual MoveList*   SelectMoveFromList(const MoveList*, int idx) = 0;
	// applies the move of `player`, all other players are assumed to submit noop; advances the player to move
	virtual GameState*  ApplyMove(const GameState*, Move*, int player) = 0;
	virtual GameState*  Next(const GameState*, const std::vector<MoveList*>& moves) = 0;
	virtual const uint32_t* GetStateHash(const GameState*) = 0;
	virtual size_t      GetStateHashSize() = 0;
	virtual string      ToString(const GameState*) = 0;
	virtual wstring     ToWString(const GameState*) = 0;
	virtual string      ToString(const Move*) = 0;
	virtual wstring     ToWString(const Move*) = 0;
	virtual void        AddRef() = 0;
	virtual void        Release() = 0;

protected:
	virtual ~IGameRules() {}
};

// Exported by a rules plugin as  extern "C" IGameRules* createGameRules(int number_of_players)
// (the reference exports the same name through BOOST_DLL_ALIAS, LinesOfActionZasady.cpp:548-551).
using CreateGameRules_t = IGameRules* (*)(int number_of_players);
