// oracle/shim/boost/property_tree/ptree.hpp -- TEST INFRASTRUCTURE ONLY.
// The reference reads player attributes with ptree::get_optional<T>(key).get_value_or(d) / get<T>(key) / put
// (MCplayer_factory.cpp:37-76); this is that surface over a string map.
#pragma once
#include <map>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
template <class T> struct optional {
	std::optional<T> v;
	explicit operator bool() const { return v.has_value(); }
	const T& get() const { return *v; }
	T get_value_or(const T& d) const { return v ? *v : d; }
};
namespace property_tree {
struct ptree {
	std::map<std::string, std::string> attrs;
	template <class T> boost::optional<T> get_optional(const std::string& key) const
	{
		boost::optional<T> o;
		auto it = attrs.find(key);
		if (it == attrs.end()) return o;
		if constexpr (std::is_same_v<T, std::string>) o.v = it->second;
		else { std::istringstream ss(it->second); T x{}; ss >> x; if (!ss.fail()) o.v = x; }
		return o;
	}
	template <class T> T get(const std::string& key) const
	{
		auto o = get_optional<T>(key);
		if (!o) throw std::runtime_error("ptree_bad_path: " + key);
		return o.get();
	}
	template <class T> void put(const std::string& key, const T& value) { std::ostringstream ss; ss << value; attrs[key] = ss.str(); }
};
}}
