// TEST INFRASTRUCTURE ONLY — the reference includes this header but uses nothing from it.
#pragma once
#include "core.hpp"
