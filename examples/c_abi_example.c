/* Minimal C host of the engine's C ABI (include/cwt_b200.h): a Morlet transform of a chirp,
 * coefficients copied back to host memory.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Lpycwt_b200 -l:libcwtb200.so -lm \
 *       -Wl,-rpath,$PWD/pycwt_b200 -o /tmp/c_abi_example && /tmp/c_abi_example
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cwt_b200.h"

int main(void) {
  const int64_t n0 = 1 << 16;
  const int n_scales = 64;
  const double dt = 1.0, two_pi = 6.283185307179586;
  double *x = malloc(sizeof(double) * (size_t)n0);
  double *scales = malloc(sizeof(double) * n_scales);
  void *W = NULL;
  cwtb_ctx *ctx = NULL;
  int rc, j;
  int64_t i;
  for (i = 0; i < n0; ++i) {
    const double t = (double)i / (double)n0;
    x[i] = sin(two_pi * (50.0 * t + (double)n0 / 8.0 * t * t));
  }
  for (j = 0; j < n_scales; ++j) scales[j] = 2.0 * pow(2.0, j / 4.0);
  if ((rc = cwtb_create(0, &ctx)) != 0) {
    fprintf(stderr, "cwtb_create failed (%d): no CUDA device?\n", rc);
    return 1;
  }
  /* page-locked result buffer: n_scales x n0 complex128 */
  if ((rc = cwtb_host_alloc(ctx, (size_t)n_scales * (size_t)n0 * 16, &W)) != 0) goto fail;
  rc = cwtb_cwt_to_host(ctx, x, /*signal_is_f32=*/0, n0, dt, scales, n_scales, CWTB_MORLET, 6.0,
                        CWTB_F64, W, /*out_f64=*/1);
  if (rc != 0) goto fail;
  printf("%s: %d scales x %lld points, %d kernel launches, %.3f ms on the device\n", cwtb_version(),
         n_scales, (long long)n0, cwtb_last_launch_count(ctx), cwtb_last_kernel_ms(ctx));
  printf("W[0][0] = %+.6e %+.6ei\n", ((double *)W)[0], ((double *)W)[1]);
  cwtb_host_free(ctx, W);
  cwtb_destroy(ctx);
  free(x);
  free(scales);
  return 0;
fail:
  fprintf(stderr, "engine error %d: %s\n", rc, cwtb_last_error(ctx));
  cwtb_destroy(ctx);
  return 1;
}
