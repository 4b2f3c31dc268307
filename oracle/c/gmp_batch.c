/*
 * Threaded GMP batch modexp: the timed CPU baseline ("the gmpy2 path").
 * TEST / BENCH INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * The reference's hot loops call third-party pow_mod / mod_inv once per element
 * (src/tno/mpc/protocols/distributed_keygen/paillier_shared_key.py:89-92,
 *  distributed_keygen.py:1094,1097); with the [gmpy] extra those are gmpy2.powmod / gmpy2.invert,
 * i.e. GMP's mpz_powm / mpz_invert.  This harness calls exactly those two functions, one element
 * at a time, from `threads` pthreads over contiguous index shards.
 *
 * gmp.h is not installed in the image, so the few prototypes used are declared here against the
 * stable GMP 6 ABI of /usr/lib/x86_64-linux-gnu/libgmp.so.10.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  int _mp_alloc;
  int _mp_size;
  unsigned long* _mp_d;
} mpz_struct;
typedef mpz_struct mpz_t[1];

void __gmpz_init(mpz_struct*);
void __gmpz_clear(mpz_struct*);
void __gmpz_import(mpz_struct*, size_t, int, size_t, int, size_t, const void*);
void* __gmpz_export(void*, size_t*, int, size_t, int, size_t, const mpz_struct*);
void __gmpz_powm(mpz_struct*, const mpz_struct*, const mpz_struct*, const mpz_struct*);
int __gmpz_invert(mpz_struct*, const mpz_struct*, const mpz_struct*);

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void export_limbs(uint32_t* dst, int L, const mpz_struct* z) {
  size_t count = 0;
  memset(dst, 0, (size_t)L * 4);
  /* least-significant word first, 4-byte words, little endian */
  __gmpz_export(dst, &count, -1, 4, -1, 0, z);
}

typedef struct {
  const uint32_t* bases;
  uint32_t* out;
  size_t begin, end;
  int L;
  const uint32_t* mod;
  int mod_limbs;
  int negative;
  const uint32_t* exp;
  int exp_limbs;
  int rc;
} fixed_job;

static void* fixed_worker(void* arg) {
  fixed_job* j = (fixed_job*)arg;
  mpz_t b, e, m, r;
  __gmpz_init(b); __gmpz_init(e); __gmpz_init(m); __gmpz_init(r);
  __gmpz_import(m, (size_t)j->mod_limbs, -1, 4, -1, 0, j->mod);
  __gmpz_import(e, (size_t)j->exp_limbs, -1, 4, -1, 0, j->exp);
  for (size_t i = j->begin; i < j->end; ++i) {
    __gmpz_import(b, (size_t)j->L, -1, 4, -1, 0, j->bases + i * (size_t)j->L);
    if (j->negative) {
      if (!__gmpz_invert(b, b, m)) { j->rc = 1; memset(j->out + i * (size_t)j->L, 0xff, (size_t)j->L * 4); continue; }
    }
    __gmpz_powm(r, b, e, m);
    export_limbs(j->out + i * (size_t)j->L, j->L, r);
  }
  __gmpz_clear(b); __gmpz_clear(e); __gmpz_clear(m); __gmpz_clear(r);
  return NULL;
}

/* out[i] = bases[i]^(+-exp) mod mod, i in [0,B); rows of L uint32 little-endian limbs. */
int gmp_powm_batch(const uint32_t* bases, uint32_t* out, size_t B, int L, const uint32_t* mod,
                   int mod_limbs, int negative, const uint32_t* exp, int exp_limbs, int threads,
                   double* seconds) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > B && B > 0) threads = (int)B;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  fixed_job* jobs = (fixed_job*)malloc(sizeof(fixed_job) * (size_t)threads);
  double t0 = now_s();
  for (int t = 0; t < threads; ++t) {
    fixed_job j = {bases, out, B * (size_t)t / (size_t)threads, B * (size_t)(t + 1) / (size_t)threads,
                   L, mod, mod_limbs, negative, exp, exp_limbs, 0};
    jobs[t] = j;
    pthread_create(&th[t], NULL, fixed_worker, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); rc |= jobs[t].rc; }
  if (seconds) *seconds = now_s() - t0;
  free(th); free(jobs);
  return rc;
}

typedef struct {
  const uint32_t* bases;
  uint32_t* out;
  size_t begin, end; /* group range */
  int K, L;
  const uint32_t* moduli;
  const uint32_t* exps;
  int exp_limbs;
} grouped_job;

static void* grouped_worker(void* arg) {
  grouped_job* j = (grouped_job*)arg;
  mpz_t b, e, m, r;
  __gmpz_init(b); __gmpz_init(e); __gmpz_init(m); __gmpz_init(r);
  for (size_t g = j->begin; g < j->end; ++g) {
    __gmpz_import(m, (size_t)j->L, -1, 4, -1, 0, j->moduli + g * (size_t)j->L);
    __gmpz_import(e, (size_t)j->exp_limbs, -1, 4, -1, 0, j->exps + g * (size_t)j->exp_limbs);
    for (int k = 0; k < j->K; ++k) {
      size_t off = (g * (size_t)j->K + (size_t)k) * (size_t)j->L;
      __gmpz_import(b, (size_t)j->L, -1, 4, -1, 0, j->bases + off);
      __gmpz_powm(r, b, e, m);
      export_limbs(j->out + off, j->L, r);
    }
  }
  __gmpz_clear(b); __gmpz_clear(e); __gmpz_clear(m); __gmpz_clear(r);
  return NULL;
}

/* out[g][k] = bases[g][k]^exps[g] mod moduli[g]: the biprimality-test batch
 * (distributed_keygen.py:1313-1329 looping :1094/:1097). */
int gmp_powm_grouped(const uint32_t* bases, uint32_t* out, size_t G, int K, int L,
                     const uint32_t* moduli, const uint32_t* exps, int exp_limbs, int threads,
                     double* seconds) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > G && G > 0) threads = (int)G;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  grouped_job* jobs = (grouped_job*)malloc(sizeof(grouped_job) * (size_t)threads);
  double t0 = now_s();
  for (int t = 0; t < threads; ++t) {
    grouped_job j = {bases, out, G * (size_t)t / (size_t)threads, G * (size_t)(t + 1) / (size_t)threads,
                     K, L, moduli, exps, exp_limbs};
    jobs[t] = j;
    pthread_create(&th[t], NULL, grouped_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  if (seconds) *seconds = now_s() - t0;
  free(th); free(jobs);
  return 0;
}
