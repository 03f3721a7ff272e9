// oracle/shim/boost/filesystem.hpp -- TEST INFRASTRUCTURE ONLY (MCTSPlayer.cpp:45 calls boost::filesystem::remove).
#pragma once
#include <filesystem>
namespace boost { namespace filesystem = std::filesystem; }
