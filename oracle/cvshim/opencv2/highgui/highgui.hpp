// TEST INFRASTRUCTURE ONLY — see opencv2/core.hpp in this directory.
// GUI calls become captures: imshow(name, m) stores a deep copy under "window:<name>"
// so the harness can read back what the reference driver displayed; imread() hands
// out the harness-registered images in call order ("imread:<n>").
#pragma once
#include "../core.hpp"
namespace cv {
enum { IMREAD_GRAYSCALE = 0 };
enum { WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1 };
enum { EVENT_LBUTTONDOWN = 1 };
typedef void (*MouseCallback)(int, int, int, int, void*);
inline int& shim_imread_counter() { static int c = 0; return c; }
inline Mat imread(const std::string&, int = IMREAD_GRAYSCALE) {
    auto it = shim_registry().find("imread:" + std::to_string(shim_imread_counter()++));
    if (it == shim_registry().end()) throw Exception("cvshim: imread called more often than images were registered");
    return it->second.clone();
}
inline void namedWindow(const std::string&, int = WINDOW_AUTOSIZE) {}
inline void resizeWindow(const std::string&, int, int) {}
inline void setMouseCallback(const std::string&, MouseCallback, void* = nullptr) {}
inline void imshow(const std::string& name, const Mat& m) { shim_registry()["window:" + name] = m.clone(); }
inline int waitKey(int = 0) { return -1; }
}  // namespace cv
