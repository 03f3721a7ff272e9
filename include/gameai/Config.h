// include/gameai/Config.h -- attribute map standing in for boost::property_tree::ptree (not installed here).
// The reference passes the XML attributes of <game .../> and <player .../> as a ptree (GamePlayer.h:9,
// GameController.h:14); only get<T>(key), get_optional<T>(key) and put(key, value) are used on it
// (MCplayer_factory.cpp:37-76, GameController.cpp:120-160), and that is what this type offers.
#pragma once
#include <map>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>

namespace gai {

template <typename T> struct Optional {
	std::optional<T> v;
	explicit operator bool() const { return v.has_value(); }
	const T& get() const { return *v; }
	T get_value_or(const T& d) const { return v ? *v : d; }
};

struct Config {
	std::map<std::string, std::string> attrs;
	std::map<std::string, Config> children;       // e.g. game/p1 .. p4 (GameLauncher.cpp:99-137)

	template <typename T> Optional<T> get_optional(const std::string& key) const
	{
		Optional<T> o;
		auto it = attrs.find(key);
		if (it == attrs.end()) return o;
		if constexpr (std::is_same_v<T, std::string>) o.v = it->second;
		else { std::istringstream ss(it->second); T x{}; ss >> x; if (!ss.fail()) o.v = x; }
		return o;
	}
	template <typename T> T get(const std::string& key) const
	{
		auto o = get_optional<T>(key);
		if (!o) throw std::runtime_error("missing config attribute: " + key);      // ptree_bad_path in the reference
		return o.get();
	}
	template <typename T> void put(const std::string& key, const T& value)
	{
		std::ostringstream ss; ss << value; attrs[key] = ss.str();
	}
};

} // namespace gai
