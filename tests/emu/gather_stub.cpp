// The NVLink exchange (csrc/gather.cu) needs CUDA IPC between processes and is not part of the
// emulated library: its kernels are exercised by tests/emu/emu_gather.cpp instead.  These
// stubs only keep the emulated library's symbol table equal to include/amsweep.h.
#include "../../include/amsweep.h"
extern "C" {
int am_gather_create(am_gather_t** out, int, int, int, uint64_t, int) { if (out) *out = nullptr; return AM_E_DEVICE; }
int am_gather_export(am_gather_t*, void*) { return AM_E_DEVICE; }
int am_gather_connect(am_gather_t*, const void*) { return AM_E_DEVICE; }
int am_gather_set_layout(am_gather_t*, const uint64_t*, const uint64_t*) { return AM_E_DEVICE; }
int am_gather_set_wire(am_gather_t*, int) { return AM_E_DEVICE; }
int am_gather_push(am_gather_t*, const void*, const void*, const void*, uint64_t, void*) { return AM_E_DEVICE; }
void* am_gather_out_idx(am_gather_t*) { return nullptr; }
void* am_gather_out_act(am_gather_t*) { return nullptr; }
void* am_gather_out_counts(am_gather_t*) { return nullptr; }
const char* am_gather_last_error(const am_gather_t*) { return "the exchange is not part of the emulated library"; }
void am_gather_destroy(am_gather_t*) {}
}
