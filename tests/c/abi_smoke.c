/* Plain-C consumer of the C ABI (no CUDA headers, no C++): loads libfpyv_b200.so, checks the ABI version and the
 * struct sizes against include/fpv_api.h as THIS translation unit sees them, and exercises the argument validation
 * (which returns before any CUDA call, so it runs on a machine without a GPU).
 *   gcc -std=c99 -I include tests/c/abi_smoke.c -ldl -o abi_smoke && ./abi_smoke fpyv_b200/libfpyv_b200.so          */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "fpv_api.h"

#define CHECK(c)                                             \
  do {                                                       \
    if (!(c)) { printf("FAILED: %s\n", #c); return 1; }      \
  } while (0)

int main(int argc, char** argv) {
  void* h = dlopen(argc > 1 ? argv[1] : "libfpyv_b200.so", RTLD_NOW);
  if (!h) { printf("dlopen: %s\n", dlerror()); return 2; }
  int (*abi)(void) = (int (*)(void))dlsym(h, "fpv_abi_version");
  int (*size_of)(int) = (int (*)(int))dlsym(h, "fpv_sizeof");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "fpv_last_error");
  int (*step)(const fpv_drone_params_t*, const fpv_drone_io_t*, void*) =
      (int (*)(const fpv_drone_params_t*, const fpv_drone_io_t*, void*))dlsym(h, "fpv_drone_step");
  int (*rollout)(const fpv_drone_params_t*, const fpv_drone_io_t*, const void*, int64_t, int32_t, uint8_t*, int64_t, void*) =
      (int (*)(const fpv_drone_params_t*, const fpv_drone_io_t*, const void*, int64_t, int32_t, uint8_t*, int64_t, void*))dlsym(h, "fpv_drone_rollout");
  CHECK(abi && size_of && last_error && step && rollout);
  CHECK(abi() == FPV_ABI_VERSION);
  CHECK(size_of(0) == (int)sizeof(fpv_drone_params_t));
  CHECK(size_of(1) == (int)sizeof(fpv_drone_io_t));
  CHECK(size_of(2) == (int)sizeof(fpv_object_t));
  CHECK(size_of(3) == (int)sizeof(fpv_stats_t));
  CHECK(size_of(4) == (int)sizeof(fpv_stick_calib_t));
  CHECK(size_of(5) == (int)sizeof(fpv_racer_params_t));
  CHECK(size_of(6) == (int)sizeof(fpv_gate_env_params_t));
  CHECK(size_of(7) == (int)sizeof(fpv_camera_params_t));
  CHECK(size_of(8) == (int)sizeof(fpv_autopilot_params_t));
  CHECK(size_of(9) == (int)sizeof(fpv_acro_params_t));
  CHECK(size_of(99) == -1);
  fpv_drone_params_t p;
  fpv_drone_io_t io;
  memset(&p, 0, sizeof p);
  memset(&io, 0, sizeof io);
  CHECK(step(0, 0, 0) == FPV_EINVAL);
  io.n = 8; io.plane_stride = 4;
  CHECK(step(&p, &io, 0) == FPV_EINVAL && strstr(last_error(), "stride"));
  io.n = 0; io.plane_stride = 0;
  CHECK(step(&p, &io, 0) == FPV_OK);                 /* empty batch: a no-op */
  CHECK(rollout(&p, &io, 0, 0, 4, 0, 0, 0) == FPV_OK);
  printf("abi_smoke: ABI %d, 10 struct layouts agree, validation ok\n", abi());
  dlclose(h);
  return 0;
}
