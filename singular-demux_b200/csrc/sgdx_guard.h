// sgdx_guard.h -- include/sgdx.h promises that no exception crosses the C boundary: every `extern "C" int` entry point
// runs its body through guarded(), which turns a throw (std::bad_alloc from a container, std::system_error from a
// thread that could not be started, ...) into an error code and the thread's sgdx_last_error text.
#pragma once
#include <exception>
#include <new>

namespace sgdx {

int fail(int code, const char *fmt, ...);

template <class F>
inline int guarded(const char *what, F &&body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return fail(-3 /* SGDX_ENOMEM */, "%s: out of host memory", what);
    } catch (const std::exception &e) {
        return fail(-3 /* SGDX_ENOMEM */, "%s: %s", what, e.what());
    } catch (...) {
        return fail(-3 /* SGDX_ENOMEM */, "%s: unknown exception", what);
    }
}

}  // namespace sgdx
