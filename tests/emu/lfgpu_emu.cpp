/* TEST-ONLY build of the host pipeline + kernels on the fiber emulator (see cuda_emu.h).
 * Produces tests/emu/liblfgpu_emu_testonly.so; nothing under lordfast_b200/ loads it. */
#include "cuda_emu.h"
#include "../../lordfast_b200/csrc/lf_pipeline.inl"

extern "C" void lf_emu_band_counts(unsigned long *ok, unsigned long *retry) { *ok = lf_emu_band_ok; *retry = lf_emu_band_retry; }
