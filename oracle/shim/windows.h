// oracle/shim/windows.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Stand-in for <windows.h> so the reference's pch.h -> framework.h -> <windows.h>
// include chain (LinesOfActionZasady/pch.h:11, framework.h:5) resolves under g++.
#pragma once
#include <cstdint>
#include <cstring>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <string>
#include <tuple>
#include <map>
#include <vector>
#include <functional>
