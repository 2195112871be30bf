// TEST INFRASTRUCTURE ONLY — the reference's driver includes matplotlibcpp.h and never uses it.
#pragma once
