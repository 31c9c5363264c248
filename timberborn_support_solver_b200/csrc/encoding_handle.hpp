// The object behind `tss_encoding*` (include/tss.h) and the process-wide instance registry that lets a solver which is handed
// a bare CNF (crates/repl/src/solver_runner.rs:8-20) find the terrain, platform set, variable map and limits it came from.
#pragma once
#include <memory>
#include <vector>

#include "host_model.hpp"

struct tss_encoding_data {
    tss::Encoding enc;
    std::vector<uint8_t> grid;
};

// a handle is a shared reference: the registry keeps encodings alive after the caller destroyed its own handle
struct tss_encoding {
    std::shared_ptr<const tss_encoding_data> d;
};

namespace tss {
// records (encoding, limits, the CNF with_limits produced); called by tss_encoding_with_limits
void instance_record(const std::shared_ptr<const tss_encoding_data>& d, const PlatformLimits& limits, const Cnf& cnf);
}  // namespace tss
