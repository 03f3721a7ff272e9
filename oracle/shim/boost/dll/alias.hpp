// oracle/shim/boost/dll/alias.hpp -- TEST INFRASTRUCTURE ONLY.
// Boost is not installed here; the oracle calls the factory function directly.
#pragma once
#define BOOST_DLL_ALIAS(func, alias)
