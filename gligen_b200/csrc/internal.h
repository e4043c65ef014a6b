// Library-internal helpers shared by the translation units of libgligen_b200.so.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include <string>

namespace glg {
// records a thread-local error message and returns -1
int set_error(const std::string& msg);
void count_launch();
// returns 0 or records the error
int check_launch(const char* what);
// bf16 tiled tensor map with 128B swizzle and zero out-of-bounds fill (host-side cache keyed by all
// arguments).  dims/box innermost first; strides in BYTES for dims 1..rank-1.  Returns 0 or records the error.
int get_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box);
int num_sms();
}  // namespace glg
