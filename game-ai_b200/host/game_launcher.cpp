// game-ai_b200/host/game_launcher.cpp -- GameLauncher + GameController for Linux: XML config -> plugins -> games.
//
// Keeps the reference's plugin model (c++/GameLauncher/GameLauncher.cpp:99-220, c++/GameController/GameController.cpp:27-217):
// the <game provider="..."> attribute names the rules plugin, each <player provider="..."> a player plugin; both are
// shared objects exporting createGameRules / createPlayer (BOOST_DLL_ALIAS in the reference, extern "C" + dlsym here);
// --p1..--p4 pick <player name=...> entries, extra "key=value" tokens override attributes (GameLauncher.cpp:119-129).
// The match loop is the reference's: every round every player is asked for a move (the one not to move returns the
// noop list), Next() advances, repeated states and the round limit abort the game (GameController.cpp:59-97).
#include "gameai/GamePlayer.h"
#include "gameai/GameRules.h"
#include "gameai/random_generator.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <fstream>
#include <iostream>
#include <mutex>
#include <thread>
#include <atomic>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>

using gai::Config;

// ---- tiny XML attribute reader: <tag a="v" b="v"> ... enough for run_config*.xml -------------------------
struct Element { std::string tag; Config attrs; };
static std::vector<Element> parse_xml(const std::string& text)
{
	std::vector<Element> out;
	size_t i = 0;
	while ((i = text.find('<', i)) != std::string::npos) {
		size_t j = text.find('>', i);
		if (j == std::string::npos) break;
		std::string body = text.substr(i + 1, j - i - 1);
		i = j + 1;
		if (body.empty() || body[0] == '?' || body[0] == '!' || body[0] == '/') continue;
		Element e; size_t k = 0;
		while (k < body.size() && !isspace((unsigned char)body[k]) && body[k] != '/') ++k;
		e.tag = body.substr(0, k);
		while (k < body.size()) {
			while (k < body.size() && (isspace((unsigned char)body[k]) || body[k] == '/')) ++k;
			size_t eq = body.find('=', k);
			if (eq == std::string::npos) break;
			std::string key = body.substr(k, eq - k);
			key.erase(std::remove_if(key.begin(), key.end(), [](char c) { return isspace((unsigned char)c); }), key.end());
			size_t q1 = body.find('"', eq); if (q1 == std::string::npos) break;
			size_t q2 = body.find('"', q1 + 1); if (q2 == std::string::npos) break;
			e.attrs.attrs[key] = body.substr(q1 + 1, q2 - q1 - 1);
			k = q2 + 1;
		}
		out.push_back(e);
	}
	return out;
}

struct Rng : IRandomGenerator {
	std::mt19937 g;
	explicit Rng(unsigned seed) : g(seed) {}
	std::vector<int> generateUniform(int lo, int hi, int n) override { std::vector<int> v(n); std::uniform_int_distribution<int> d(lo, hi); for (auto& x : v) x = d(g); return v; }
	void release() override { delete this; }
};

// `dirs`: colon-separated list of directories searched in order (the launcher's own directory first, then --plugins)
static void* load_plugin(const std::string& provider, const std::string& dirs)
{
	std::string lower = provider; std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
	std::vector<std::string> names;
	std::istringstream ds(dirs); std::string dir;
	while (std::getline(ds, dir, ':')) if (!dir.empty()) { names.push_back(dir + "/lib" + provider + ".so"); names.push_back(dir + "/lib" + lower + ".so"); }
	names.push_back("lib" + provider + ".so"); names.push_back("lib" + lower + ".so");
	for (const std::string& nm : names) if (void* h = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL)) return h;
	std::fprintf(stderr, "cannot load plugin '%s': %s\n", provider.c_str(), dlerror());
	return nullptr;
}

enum class SingleGameResult { Win = 0, RoundLimit = 1, StateLoop = 2 };

// GameController.cpp:27-115
static std::string g_first_move;      // text of the first move made in the last game (used by --rl 1 probes)
static std::mutex g_io;                // stdout lines and g_first_move (several controller threads)
static SingleGameResult run_single_game(IGameRules* rules, std::vector<IGamePlayer*>& players, GameState* state, int round_limit, int score[], int verbose, int* rounds_out)
{
	const int np = (int)players.size();
	std::vector<GameState*> pks(np);
	for (int i = 0; i < np; ++i) { pks[i] = rules->CreatePlayerKnownState(state, i); players[i]->startNewGame(pks[i]); }
	std::set<std::string> seen;
	SingleGameResult result = SingleGameResult::Win;
	std::vector<MoveList*> moves(np);
	int round = 0;
	for (;;) {
		for (int i = 0; i < np; ++i) moves[i] = players[i]->selectMove(pks[i]);
		if (round == 0) {
			const int cp0 = rules->GetCurrentPlayer(state);
			if (rules->GetNumMoves(moves[cp0]) > 0) { auto [mv0, p0] = rules->GetMoveFromList(moves[cp0], 0); (void)p0; std::lock_guard<std::mutex> lk(g_io); g_first_move = rules->ToString(mv0); }
		}
		if (verbose) {
			const int cp = rules->GetCurrentPlayer(state);
			auto [mv, p] = rules->GetMoveFromList(moves[cp], 0); (void)p;
			std::printf("round %d player %d (%s): %s\n", round, cp, players[cp]->getName().c_str(), rules->ToString(mv).c_str());
		}
		state = rules->Next(state, moves);
		const std::string key = rules->ToString(state);
		const bool loop = !seen.insert(key).second;
		for (int i = 0; i < np; ++i) rules->UpdatePlayerKnownState(pks[i], state, moves);
		for (int i = 0; i < np; ++i) rules->ReleaseMoveList(moves[i]);
		++round;
		if (loop) { result = SingleGameResult::StateLoop; break; }
		if (round >= round_limit) { result = SingleGameResult::RoundLimit; break; }
		if (rules->IsTerminal(state)) break;
	}
	rules->Score(state, score);
	if (verbose) std::printf("final state:\n%s\n", rules->ToString(state).c_str());
	for (int i = 0; i < np; ++i) {
		const GameResult gr = result == SingleGameResult::Win ? (score[i] == 0 ? GameResult::Loose : (np == 2 && score[i] == 50) ? GameResult::Draw : GameResult::Win)
		                    : result == SingleGameResult::RoundLimit ? GameResult::AbortedByRoundLimit : GameResult::AbortedByStateLoop;
		players[i]->endGame(score[i], gr);
		rules->ReleaseGameState(pks[i]);
	}
	rules->ReleaseGameState(state);
	*rounds_out = round;
	return result;
}

// mergeResults (common/api/GameController.h:41): histograms and averages add up, numbers add, text keeps the last value
static void merge_metrics(NamedMetrics_t& into, const NamedMetrics_t& from)
{
	for (auto& kv : from) {
		auto it = into.find(kv.first);
		if (it == into.end()) { into[kv.first] = kv.second; continue; }
		if (auto h = std::get_if<Histogram<long>>(&kv.second)) { if (auto d = std::get_if<Histogram<long>>(&it->second)) *d += *h; }
		else if (auto a = std::get_if<Average>(&kv.second)) { if (auto d = std::get_if<Average>(&it->second)) *d += *a; }
		else if (auto x = std::get_if<double>(&kv.second)) { if (auto d = std::get_if<double>(&it->second)) *d += *x; }
		else it->second = kv.second;
	}
}
static std::string xml_escape(const std::string& s)
{
	std::string o;
	for (char c : s) { if (c == '&') o += "&amp;"; else if (c == '<') o += "&lt;"; else if (c == '>') o += "&gt;"; else if (c == '"') o += "&quot;"; else o += c; }
	return o;
}

// a crash inside a plugin should say where: print the call stack before dying
static void crash_handler(int sig)
{
	void* frames[64];
	const int n = backtrace(frames, 64);
	const char msg[] = "GameLauncher: fatal signal, call stack:\n";
	if (write(2, msg, sizeof msg - 1) < 0) {}
	backtrace_symbols_fd(frames, n, 2);
	signal(sig, SIG_DFL); raise(sig);
}

int main(int argc, char** argv)
{
	signal(SIGSEGV, crash_handler); signal(SIGABRT, crash_handler); signal(SIGFPE, crash_handler);
	std::string xml = "run_config_loa.xml", plugin_dir, save_file;
	std::vector<std::string> pnames(4);
	int num_games = -1, round_limit = -1, verbose = 0, num_threads = -1;
	bool progress = false, quiet = false;
	std::string start_hash;
	unsigned seed = 12345;
	{	// plugins live next to the executable
		std::string self = argv[0]; size_t sl = self.find_last_of('/'); plugin_dir = sl == std::string::npos ? "." : self.substr(0, sl);
	}
	for (int i = 1; i < argc; ++i) {
		std::string a = argv[i];
		auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
		if (a == "--xml") xml = next();
		else if (a == "--p1" || a == "--p2" || a == "--p3" || a == "--p4") pnames[a[3] - '1'] = next();
		else if (a == "--ng") num_games = atoi(next().c_str());
		else if (a == "--rl") round_limit = atoi(next().c_str());
		else if (a == "--threads" || a == "-t") num_threads = atoi(next().c_str());           // GameLauncher.cpp:152
		else if (a == "--progress" || a == "-p") progress = true;                           // :145
		else if (a == "--quiet" || a == "-q") quiet = true;                                 // :153
		else if (a == "--save") save_file = next();
		else if (a == "--seed") seed = (unsigned)strtoul(next().c_str(), nullptr, 10);
		else if (a == "--verbose" || a == "-v") verbose = 1;
		else if (a == "--plugins") plugin_dir += ":" + next();                              // extra plugin directories (colon separated)
		else if (a == "--start") start_hash = next();      // "white,black,cp" as hex/decimal words: play from this position
		else if (a == "--help" || a == "-h") { std::printf("usage: GameLauncher --xml cfg.xml --p1 \"name [key=value ...]\" --p2 name [--ng N] [--rl N] [--threads N] [--seed S] [--save results.xml] [-v] [-p] [-q]\n"); return 0; }
	}
	std::ifstream f(xml);
	if (!f) { std::fprintf(stderr, "cannot open %s\n", xml.c_str()); return 2; }
	std::stringstream ss; ss << f.rdbuf();
	auto elems = parse_xml(ss.str());
	Config game; std::vector<Config> player_defs;
	for (auto& e : elems) { if (e.tag == "game") game = e.attrs; else if (e.tag == "player") player_defs.push_back(e.attrs); }
	if (num_games < 0) num_games = game.get_optional<int>("num_games").get_value_or(1);
	if (round_limit < 0) round_limit = game.get_optional<int>("round_limit").get_value_or(300);
	if (num_threads < 0) num_threads = game.get_optional<int>("num_threads").get_value_or(1);       // GameController.cpp:122
	num_threads = std::max(1, std::min(num_threads, std::max(1, num_games)));
	if (save_file.empty()) save_file = game.get_optional<std::string>("save").get_value_or("");      // run_config_loa.xml:3 save="results.xml"
	if (!save_file.empty() && save_file.find('/') == std::string::npos) {
		const std::string od = game.get_optional<std::string>("out_dir").get_value_or("");
		if (!od.empty() && od.find('\\') == std::string::npos) save_file = od + "/" + save_file;     // a Windows path from the reference's XML is not a directory here
	}

	void* hr = load_plugin(game.get<std::string>("provider"), plugin_dir);
	if (!hr) return 3;
	auto create_rules = (CreateGameRules_t)dlsym(hr, "createGameRules");
	if (!create_rules) { std::fprintf(stderr, "plugin lacks createGameRules\n"); return 3; }

	// selectPlayerConfigs (GameLauncher.cpp:99-137): "--p1 'mcts key=value ...'"
	std::vector<Config> pcfg;
	for (auto& spec : pnames) {
		if (spec.empty()) continue;
		std::istringstream ts(spec); std::string name; ts >> name;
		Config c; bool found = false;
		for (auto& d : player_defs) if (d.get_optional<std::string>("name").get_value_or("") == name) { c = d; found = true; break; }
		if (!found) { std::fprintf(stderr, "no <player name=\"%s\"> in %s\n", name.c_str(), xml.c_str()); return 2; }
		std::string tok, full = name;
		while (ts >> tok) { size_t eq = tok.find('='); if (eq != std::string::npos) { c.attrs[tok.substr(0, eq)] = tok.substr(eq + 1); full += "_" + tok; } }
		c.attrs["fullname"] = full;
		pcfg.push_back(c);
	}
	const int np = (int)pcfg.size();
	if (np < 2) { std::fprintf(stderr, "need at least --p1 and --p2\n"); return 2; }
	std::vector<CreatePlayer_t> factories;
	for (int i = 0; i < np; ++i) {
		pcfg[i].put("number_of_players", np);                                        // GameController.cpp:152-155
		void* hp = load_plugin(pcfg[i].get<std::string>("provider"), plugin_dir);
		if (!hp) return 3;
		auto create_player = (CreatePlayer_t)dlsym(hp, "createPlayer");
		if (!create_player) { std::fprintf(stderr, "plugin lacks createPlayer\n"); return 3; }
		factories.push_back(create_player);
	}

	// results of the whole run (merged under a mutex, GameController.cpp:208-210)
	std::mutex mtx;
	std::vector<double> pts(np, 0); std::vector<int> wins(np, 0), loses(np, 0);
	std::vector<NamedMetrics_t> stats((size_t)np); std::vector<std::string> names((size_t)np);
	std::map<std::string, int> game_results;
	int aborted = 0; long rounds_total = 0; int failed = 0; std::string failure;
	std::atomic<int> games_done{ 0 };
	const auto t0 = std::chrono::steady_clock::now();

	// one controller thread = its own rules object, players (a GPU player: its own gai_ctx) and generator, never shared
	// (GameController.cpp:167-204); the games are split evenly, the first threads take the remainder
	auto worker = [&](int instance) {
		const int my_games = num_games / num_threads + (instance < num_games % num_threads ? 1 : 0);
		IGameRules* rules = create_rules(np);
		Rng* rng = new Rng(seed + 7919u * (unsigned)instance);
		std::vector<IGamePlayer*> players;
		std::vector<double> t_pts(np, 0); std::vector<int> t_wins(np, 0), t_loses(np, 0);
		std::map<std::string, int> t_results; int t_aborted = 0; long t_rounds = 0;
		try {
			for (int i = 0; i < np; ++i) {
				Config c = pcfg[i];
				if (!c.get_optional<unsigned>("random_seed")) c.put("random_seed", seed + 1000u * (unsigned)(i + 1) + 7919u * (unsigned)instance);
				else if (instance > 0) c.put("random_seed", c.get<unsigned>("random_seed") + 7919u * (unsigned)instance);      // threads must not replay each other's games
				IGamePlayer* p = factories[i](i, &c);
				p->setGameRules(rules);
				players.push_back(p);
			}
			for (int g = 0; g < my_games; ++g) {
				int score[4] = { 0, 0, 0, 0 }, rounds = 0;
				GameState* st;
				if (start_hash.empty()) st = rules->CreateRandomInitialState(rng);
				else {
					uint64_t wds[3] = { 0, 0, 0 }; char* e = nullptr; const char* q = start_hash.c_str();
					for (int k = 0; k < 3; ++k) { wds[k] = strtoull(q, &e, 0); q = (*e == ',') ? e + 1 : e; }
					st = rules->CreateInitialStateFromHash(reinterpret_cast<const uint32_t*>(wds));       // LOA: the 24-byte state is its own hash
				}
				const SingleGameResult r = run_single_game(rules, players, st, round_limit, score, verbose && num_threads == 1, &rounds);
				t_rounds += rounds;
				if (r != SingleGameResult::Win) ++t_aborted;
				t_results[r == SingleGameResult::Win ? "Win" : r == SingleGameResult::RoundLimit ? "RoundLimit" : "StateLoop"] += 1;       // singleRunResultString
				for (int i = 0; i < np; ++i) { t_pts[i] += score[i]; if (r == SingleGameResult::Win) { if (score[i] == 0) ++t_loses[i]; else if (!(np == 2 && score[i] == 50)) ++t_wins[i]; } }
				const int done = ++games_done;
				std::lock_guard<std::mutex> lk(g_io);
				if (!quiet) {
					std::printf("game %d: %s after %d rounds, score", done, r == SingleGameResult::Win ? "finished" : r == SingleGameResult::RoundLimit ? "round limit" : "state loop", rounds);
					for (int i = 0; i < np; ++i) std::printf(" %s=%d", players[i]->getName().c_str(), score[i]);
					std::printf("\n"); std::fflush(stdout);
				}
				if (progress) {                                                               // createProgressBar(total_progress, number_of_games, type)
					const int w = 40, fill = (int)((long)w * done / std::max(1, num_games));
					std::fprintf(stderr, "\r[%.*s%*s] %d/%d", fill, "########################################", w - fill, "", done, num_games);
					if (done == num_games) std::fprintf(stderr, "\n");
					std::fflush(stderr);
				}
			}
		} catch (const char* msg) { std::lock_guard<std::mutex> lk(mtx); ++failed; failure = msg; }
		  catch (const std::exception& e) { std::lock_guard<std::mutex> lk(mtx); ++failed; failure = e.what(); }
		{
			std::lock_guard<std::mutex> lk(mtx);
			for (int i = 0; i < np && i < (int)players.size(); ++i) {
				pts[i] += t_pts[i]; wins[i] += t_wins[i]; loses[i] += t_loses[i];
				merge_metrics(stats[(size_t)i], players[i]->getGameStats());               // appendPlayerStats
				names[(size_t)i] = players[i]->getName();
			}
			for (auto& kv : t_results) game_results[kv.first] += kv.second;
			aborted += t_aborted; rounds_total += t_rounds;
		}
		for (auto* p : players) p->release();
		rules->Release();
		rng->release();
	};
	if (num_threads == 1) worker(0);
	else {
		std::vector<std::thread> th;
		for (int i = 0; i < num_threads; ++i) th.emplace_back(worker, i);
		for (auto& t : th) t.join();
	}
	if (failed) { std::fprintf(stderr, "error: %s\n", failure.c_str()); return 4; }
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

	// results.xml (saveXmlResults; the reference stores pN.pts / pN.win / pN.lose, num_rounds, num_games, game_results and the
	// players' metrics as attributes, GameController.cpp:101-111,199,213-215)
	if (!save_file.empty()) {
		std::ofstream o(save_file);
		if (!o) std::fprintf(stderr, "cannot write %s\n", save_file.c_str());
		else {
			o << "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n<results num_games=\"" << num_games << "\" num_rounds=\"" << rounds_total << "\" num_threads=\"" << num_threads
			  << "\" time_s=\"" << secs << "\" game_results=\"";
			bool first = true;
			for (auto& kv : game_results) { o << (first ? "" : ",") << kv.first << ":" << kv.second; first = false; }
			o << "\"";
			for (int i = 0; i < np; ++i) o << " p" << i + 1 << ".pts=\"" << pts[i] << "\" p" << i + 1 << ".win=\"" << wins[i] << "\" p" << i + 1 << ".lose=\"" << loses[i] << "\"";
			o << ">\n";
			for (int i = 0; i < np; ++i) {
				o << "  <p" << i + 1 << " name=\"" << xml_escape(names[(size_t)i]) << "\"";
				for (auto& kv : stats[(size_t)i]) o << " " << kv.first << "=\"" << xml_escape(metric_to_string(kv.second)) << "\"";
				o << " />\n";
			}
			o << "</results>\n";
		}
	}

	std::printf("{\"games\": %d, \"aborted\": %d, \"rounds\": %ld, \"seconds\": %.3f, \"threads\": %d, \"first_move\": \"%s\", \"players\": [", num_games, aborted, rounds_total, secs, num_threads, g_first_move.c_str());
	for (int i = 0; i < np; ++i) {
		std::printf("%s{\"name\": \"%s\", \"pts\": %.1f, \"win\": %d, \"lose\": %d, \"stats\": {", i ? ", " : "", names[(size_t)i].c_str(), pts[i], wins[i], loses[i]);
		bool first = true;
		for (auto& kv : stats[(size_t)i]) { std::printf("%s\"%s\": \"%s\"", first ? "" : ", ", kv.first.c_str(), metric_to_string(kv.second).c_str()); first = false; }
		std::printf("}}");
	}
	std::printf("]}\n");
	return 0;
}
