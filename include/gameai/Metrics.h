// include/gameai/Metrics.h -- player statistics, the subset of c++/common/api/Metrics.h the MCTS player reports
// (MCTSPlayer.cpp:149-165: histograms "runs_per_move", "branching_factor", "path_size", averages of ratios).
#pragma once
#include <map>
#include <sstream>
#include <string>
#include <variant>

template <typename T>
struct Histogram {
	T bucket = 1;
	std::map<T, unsigned> values;
	Histogram& Bucket(T b) { bucket = b; return *this; }
	void insert(T v) { values[(v / bucket) * bucket] += 1; }
	Histogram& operator+=(const Histogram& o) { for (auto& kv : o.values) values[kv.first] += kv.second; return *this; }
	std::string to_string() const
	{
		std::ostringstream ss; bool first = true;
		for (auto& kv : values) { if (!first) ss << ","; ss << kv.first << ":" << kv.second; first = false; }
		return ss.str();
	}
};
struct Average {
	double sum = 0; unsigned long n = 0;
	Average() {}
	explicit Average(double v) : sum(v), n(1) {}
	void insert(double v) { sum += v; ++n; }
	double value() const { return n ? sum / n : 0.0; }
	Average& operator+=(const Average& o) { sum += o.sum; n += o.n; return *this; }
	std::string to_string() const { std::ostringstream ss; ss << value(); return ss.str(); }
};
using Metric_t = std::variant<Histogram<long>, Average, double, std::string>;
using NamedMetrics_t = std::map<std::string, Metric_t>;

inline std::string metric_to_string(const Metric_t& m)
{
	if (auto h = std::get_if<Histogram<long>>(&m)) return h->to_string();
	if (auto a = std::get_if<Average>(&m)) return a->to_string();
	if (auto d = std::get_if<double>(&m)) { std::ostringstream ss; ss << *d; return ss.str(); }
	return std::get<std::string>(m);
}
