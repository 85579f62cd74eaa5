/* oracle/ref_counters.hpp -- TEST INFRASTRUCTURE.
 * Per-thread traversal counters for the instrumented reference build (oracle/instrument.sh
 * inserts WPT_REF_COUNT(...) lines into generated copies of bvh.hpp / wurblpt.hpp). They give
 * the reference's own node visits and leaf tests per closest-hit query, the n and l of
 * SURVEY.md section 8(d)'s algorithmic-bytes formula. */
#pragma once

#include <omp.h>
#include <cstdint>

#define WPT_REF_MAX_THREADS 1024

struct alignas(64) WptRefCounters {
    uint64_t rays = 0;      /* closest-hit queries issued by tracePath() */
    uint64_t calls = 0;     /* all BVH::hit() calls, including nested medium boundaries */
    uint64_t nodes = 0;     /* BVHNodeLinear visits */
    uint64_t leaves = 0;    /* Hitable::hit() calls made from BVH leaves */
};

extern WptRefCounters wpt_ref_counters[WPT_REF_MAX_THREADS];

#define WPT_REF_COUNT(what) (wpt_ref_counters[omp_get_thread_num() % WPT_REF_MAX_THREADS].what++)
