/* Compares hs_sinf/hs_cosf/hs_expf (tfe-slam-ucl_b200/csrc/hs_libm_exact.h, host build) with the host
 * libm over a strided sweep of all float bit patterns in the supported domain.
 * usage: libm_exact_check <stride> [threads]   (stride 1 = exhaustive)
 * prints "<fn> checked <n> mismatches <m>" per function; exit code 0 iff no mismatches. */
#include "hs_libm_exact.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

typedef struct { uint32_t stride, tid, threads; unsigned long long n[3], bad[3]; } arg_t;

static float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void* worker(void* vp) {
  arg_t* a = (arg_t*)vp;
  /* sin/cos: |x| < 120 -> bit patterns [0, asuint(120.0f)) for both signs */
  uint32_t lim = hs_asuint(120.0f);
  for (uint64_t u = (uint64_t)a->tid * a->stride; u < lim; u += (uint64_t)a->stride * a->threads) {
    for (int sgn = 0; sgn < 2; ++sgn) {
      float x = asfloat((uint32_t)u | (sgn ? 0x80000000u : 0u));
      float r0 = sinf(x), r1 = hs_sinf(x);
      a->n[0]++; if (hs_asuint(r0) != hs_asuint(r1)) { if (a->bad[0]++ < 3) fprintf(stderr, "sinf(%a) libm %a ours %a\n", x, r0, r1); }
      r0 = cosf(x); r1 = hs_cosf(x);
      a->n[1]++; if (hs_asuint(r0) != hs_asuint(r1)) { if (a->bad[1]++ < 3) fprintf(stderr, "cosf(%a) libm %a ours %a\n", x, r0, r1); }
    }
  }
  /* expf: every finite float and the infinities */
  uint32_t lim2 = 0x7f800000u;
  for (uint64_t u = (uint64_t)a->tid * a->stride; u <= lim2; u += (uint64_t)a->stride * a->threads) {
    for (int sgn = 0; sgn < 2; ++sgn) {
      float x = asfloat((uint32_t)u | (sgn ? 0x80000000u : 0u));
      float r0 = expf(x), r1 = hs_expf(x);
      a->n[2]++; if (hs_asuint(r0) != hs_asuint(r1)) { if (a->bad[2]++ < 3) fprintf(stderr, "expf(%a) libm %a ours %a\n", x, r0, r1); }
    }
  }
  return NULL;
}

int main(int argc, char** argv) {
  uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], NULL, 10) : 1009;
  int threads = argc > 2 ? atoi(argv[2]) : 8;
  if (stride < 1) stride = 1;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  arg_t args[256]; pthread_t th[256];
  for (int t = 0; t < threads; ++t) { memset(&args[t], 0, sizeof(arg_t)); args[t].stride = stride; args[t].tid = t; args[t].threads = threads; pthread_create(&th[t], NULL, worker, &args[t]); }
  unsigned long long n[3] = {0, 0, 0}, bad[3] = {0, 0, 0};
  for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); for (int k = 0; k < 3; ++k) { n[k] += args[t].n[k]; bad[k] += args[t].bad[k]; } }
  const char* names[3] = {"sinf", "cosf", "expf"};
  for (int k = 0; k < 3; ++k) printf("%s checked %llu mismatches %llu\n", names[k], n[k], bad[k]);
  return (bad[0] | bad[1] | bad[2]) ? 1 : 0;
}
