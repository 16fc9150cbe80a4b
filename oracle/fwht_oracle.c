/* CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * C restatement of the reference's Walsh-Hadamard butterfly:
 *   /root/reference/rla/srht.py:14-36  (_fht_1d: stages h = 1, 2, 4, ...,
 *                                       (x, y) -> (x + y, x - y))
 *   /root/reference/rla/srht.py:93-96  (_fht_2d_parallel: one row per thread)
 * The 2**(d/2) normalisation (srht.py:36) is applied by the Python caller so
 * that it is literally NumPy's `a /= 2**(d/2)`.
 *
 * Threads: plain pthreads pulling rows from a shared counter (the reference's
 * numba `prange` over rows), so no OpenMP runtime is needed.
 * Built by oracle/Makefile into oracle/_build/libfwht_oracle.so.
 */
#include <pthread.h>
#include <stdint.h>
#include <unistd.h>

static void wht_row(double *a, int64_t n)
{
    for (int64_t h = 1; h < n; h <<= 1) {
        for (int64_t i = 0; i < n; i += h << 1) {
            double *lo = a + i, *hi = a + i + h;
            for (int64_t j = 0; j < h; ++j) {
                double x = lo[j], y = hi[j];
                lo[j] = x + y;
                hi[j] = x - y;
            }
        }
    }
}

int oracle_max_threads(void)
{
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

typedef struct {
    double *a;
    int64_t m, n;
    int64_t next;          /* next unclaimed row, protected by mu */
    pthread_mutex_t mu;
} job_t;

static void *worker(void *p)
{
    job_t *job = (job_t *)p;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        int64_t r = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (r >= job->m) break;
        wht_row(job->a + r * job->n, job->n);
    }
    return 0;
}

/* Unnormalised in-place WHT of each of the m rows (length n = 2**d, contiguous).
 * nthreads <= 0 means "all online cores". */
void oracle_fwht_rows_f64(double *a, int64_t m, int64_t n, int nthreads)
{
    if (nthreads <= 0) nthreads = oracle_max_threads();
    if (nthreads > m) nthreads = (int)m;
    if (nthreads <= 1) {
        for (int64_t r = 0; r < m; ++r) wht_row(a + r * n, n);
        return;
    }
    if (nthreads > 256) nthreads = 256;
    job_t job = {a, m, n, 0, PTHREAD_MUTEX_INITIALIZER};
    pthread_t tid[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; ++t)
        if (pthread_create(&tid[started], 0, worker, &job) == 0) ++started;
    worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], 0);
}
