// Exceptions.h — the reference's exception types (upstream include/Exceptions.h:4-10, plus the
// MeshError / ParserError of its older API) and the mapping from C-ABI status codes.
#pragma once
#include <stdexcept>
#include <string>

#include "../swe_b200.h"

struct DomainError : std::runtime_error { using std::runtime_error::runtime_error; };
struct SolverError : std::runtime_error { using std::runtime_error::runtime_error; };
struct MeshError : std::runtime_error { using std::runtime_error::runtime_error; };
struct DeviceError : std::runtime_error { using std::runtime_error::runtime_error; };

namespace swe_detail {
inline void check(int status, const swe_ctx *ctx = nullptr) {
    if (status == SWE_OK) return;
    const char *m = swe_last_error(ctx);
    const std::string msg = m ? m : "unknown error";
    switch (status) {
        case SWE_ERR_INVALID: throw DomainError(msg);
        case SWE_ERR_IO: throw MeshError(msg);
        case SWE_ERR_NUMERIC: throw SolverError(msg);
        default: throw DeviceError(msg);  // CUDA failure / no device: there is no CPU fallback
    }
}
}  // namespace swe_detail
