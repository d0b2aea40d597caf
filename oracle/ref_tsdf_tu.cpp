// TEST INFRASTRUCTURE ONLY (oracle/). Translation unit that compiles the reference's
// src/chad/tsdf.cpp UNMODIFIED (included from /root/reference via -I, never copied). With
// -DCHAD_REF_STABLE the token `sort` is renamed to `stable_sort` while the reference is
// parsed (morton.hpp:89) -- the tie-break canonicalisation of SURVEY.md section 8c. The
// standard headers are included first so the macro cannot touch them.
#include <algorithm>
#include <array>
#include <bit>
#include <bitset>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>
#include <fmt/base.h>
#include <glm/glm.hpp>
#include <gtl/phmap.hpp>
#include <libmorton/morton.h>
#ifdef CHAD_REF_STABLE
#define sort stable_sort
#endif
#include "src/chad/tsdf.cpp"
