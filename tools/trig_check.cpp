// tools/trig_check.cpp -- csrc/trig.cuh (host build) against the running libm's sinf / cosf on EVERY float below a limit.
//   g++ -O2 -ffp-contract=off -mfma -pthread -o trig_check tools/trig_check.cpp -lm ;  ./trig_check [limit = 16] [threads = 16]
// Prints "mismatches 0 0" and exits 0 when the restatement is bit-identical for both functions and both signs.
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "../hardware-acceleration-of-lidar-slam_b200/csrc/trig.cuh"

struct Job { uint32_t lo, hi; long bad_s, bad_c; uint32_t first; };

static void *run(void *a)
{
    Job *j = (Job *)a;
    for (uint32_t u = j->lo; u < j->hi; ++u)
        for (int sg = 0; sg < 2; ++sg) {
            const uint32_t v = u | (sg ? 0x80000000u : 0u);
            float y;
            memcpy(&y, &v, 4);
            const float s0 = sinf(y), s1 = glibc_trig::sincos(y, 0), c0 = cosf(y), c1 = glibc_trig::sincos(y, 1);
            if (memcmp(&s0, &s1, 4)) { if (!j->bad_s && !j->bad_c) j->first = v; j->bad_s++; }
            if (memcmp(&c0, &c1, 4)) { if (!j->bad_s && !j->bad_c) j->first = v; j->bad_c++; }
        }
    return 0;
}

int main(int argc, char **argv)
{
    const float lim = argc > 1 ? (float)atof(argv[1]) : glibc_trig::MAX_ABS;
    int nt = argc > 2 ? atoi(argv[2]) : 16;
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    uint32_t top;
    memcpy(&top, &lim, 4);
    top += 1;                                   // the limit itself included
    pthread_t th[64];
    Job jobs[64];
    for (int t = 0; t < nt; ++t) {
        jobs[t].lo = (uint32_t)((uint64_t)top * t / nt);
        jobs[t].hi = (uint32_t)((uint64_t)top * (t + 1) / nt);
        jobs[t].bad_s = jobs[t].bad_c = 0;
        jobs[t].first = 0;
        pthread_create(&th[t], 0, run, &jobs[t]);
    }
    long bs = 0, bc = 0;
    for (int t = 0; t < nt; ++t) {
        pthread_join(th[t], 0);
        bs += jobs[t].bad_s;
        bc += jobs[t].bad_c;
        if (jobs[t].bad_s || jobs[t].bad_c) printf("first mismatch in slice %d: 0x%08x\n", t, jobs[t].first);
    }
    printf("|y| <= %g: %u floats x 2 signs x {sinf, cosf}: mismatches %ld %ld\n", lim, top, bs, bc);
    return bs || bc ? 1 : 0;
}
