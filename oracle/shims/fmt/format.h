// TEST INFRASTRUCTURE ONLY (oracle/): see base.h in this directory.
#pragma once
#include "base.h"
