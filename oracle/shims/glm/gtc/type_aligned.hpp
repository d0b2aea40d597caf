// TEST INFRASTRUCTURE ONLY (oracle/): see glm.hpp in this directory.
#pragma once
#include "../glm.hpp"
