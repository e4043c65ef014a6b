// Library-internal helpers shared by the translation units of libgligen_b200.so.
#pragma once
#include <string>

namespace glg {
// records a thread-local error message and returns -1
int set_error(const std::string& msg);
void count_launch();
// returns 0 or records the error
int check_launch(const char* what);
}  // namespace glg
