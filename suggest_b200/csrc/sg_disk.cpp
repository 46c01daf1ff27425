// sg_disk.cpp — reader for indexes written by the reference's `suggest indexer` (placeholder until the
// gob / VB / skipping / roaring decoders land; see DESIGN.md "next rows").
#include "sg_host.h"

namespace sg {

std::string build_from_disk(HostIndex *, const char *, const char *) {
    return "on-disk index reader is not part of this build yet";
}

}  // namespace sg
