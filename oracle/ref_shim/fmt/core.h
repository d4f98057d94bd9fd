/* oracle/ref_shim: the reference only needs fmt::format to build exception / toString text, never on a numeric path. TEST INFRASTRUCTURE. */
#pragma once
#include <string>
namespace fmt { template <typename... A> std::string format(const char *f, const A &...) { return std::string(f); }
                template <typename... A> std::string format(const std::string &f, const A &...) { return f; }
                template <typename... A> void print(const char *, const A &...) {} }
