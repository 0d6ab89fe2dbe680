// cr_rcpar.cuh -- the range recurrence of one coder stream, cut into jobs that run side by side (rc_variant 8).
//
// Replaces the serial walk of `range` in range_encoder_encode (src/cr-rangecoder.c:60-70: range /= sum; range *= frq; renormalise)
// for a whole (block, stream).  The recurrence  r' = norm(floor(r / sum) * frq)  cannot be speculated from a guessed state, but it
// FORGETS: floor(r / sum) merges every r of one quotient bucket, so the image of ALL 2^32 - 2^24 possible states shrinks to a few
// hundred values within some thousand symbols (measured on the text workload: ~60000 / sqrt(symbols), profiles/round2_summary.md).
// That makes the exit state of a piece of the stream a function of its entry state with a small, enumerable range:
//
//   plan    the stream is cut into jobs of ~T symbols; job c >= 1 starts at the symbol with the LARGEST sum near its nominal start
//           (a_c): behind that symbol only  2^32 / sum  states are possible, whatever came before.
//   seed    (A1) job c enumerates those states implicitly (q = qlo .. qhi), runs S symbols on each, drops equal neighbours.
//   track   (A2, A3) the surviving states are stepped in registers to the start of the next job, merged values dropped as they
//           appear: E_c = every state the coder can possibly be in at a_(c+1).
//   follow  (B) job c steps every candidate of E_(c-1) through its own symbols: F_c[j] = where candidate j ends up.
//   resolve (C) one walk over the jobs of a stream: entry_(c+1) = F_c[ index of entry_c in E_(c-1) ].
//   emit    (D) every job re-walks its symbols from its true entry state and writes the quotient and the top-bit index of every
//           symbol (what k_range_chain<7> writes), then CHECKS that it arrives at the entry state of the next job.
//
// Exactness does not rest on the set tracking: phase D is the same step function as the serial chain, and a stream whose jobs do
// not link up (or whose seeds outgrow their buffers) is flagged and re-done by the serial walk (k_rcp_emit in whole-stream mode).
// Streams whose sums are all small (BMP data through the LZP coder: sums below 2^14) never shrink far enough; their jobs are
// merged at plan time and the stream runs as one serial job, i.e. exactly as before.
//
// Arithmetic: the double-precision form of the step (rc_dp_step, cr_rc.cuh): two dependent DFMA and three integer operations per
// state and symbol, states kept as doubles with exact integer values in [2^24, 2^32).
#pragma once
#include "cr_rc.cuh"

#ifndef CRGPU_SIM
#define RCP_MMAX      262144u      // most states a seed may enumerate (2^32 / sum at the seed symbol: sum >= 2^14)
#define RCP_CAP_A     16384u       // most states a seed may hand to the tracking kernels
#define RCP_CAP_E     2048u        // most states of an exit set E_c
#define RCP_S0        64u          // symbols a seed runs before it counts its survivors (doubled while they exceed RCP_CAP_A)
#define RCP_SMAX      512u
#define RCP_WINDOW    2048u        // a job's start is the largest sum among this many symbols behind its nominal start
#define RCP_NONE      0xFFFFFFFFu

enum { RCP_LIVE = 1, RCP_MERGED = 2 };
struct RcpStream { unsigned long long i0, i1; uint32_t first_job, njobs, is_main, flags; };   // flags != 0: redo serially
struct RcpJob {
    unsigned long long nominal, a, pos, end;  // nominal start; real start (seed symbol; stream start for job 0); next symbol of the tracked set; end (k_rcp_links)
    uint32_t stream, status, count, ecount;   // count: states handed from A1 to A2/A3; ecount: |E_c|
    uint32_t next, prev;                      // neighbouring live jobs of the stream (k_rcp_links)
    double entry;                             // true state on entry (phase C)
};
struct RcpStats { unsigned long long state_steps; uint32_t live_jobs, merged_jobs, seed_retries, demoted_jobs, flagged_streams, max_e; unsigned long long phase_steps[4];   // phase_steps: seed, mid, late, follow
                  uint32_t decided_serial, pilot_jobs; unsigned long long est_steps, serial_equiv; };

// The cut pays only while the tracked sets are small: the tracking passes step every possible state through every symbol (twice), and
// their cost does not shrink with the number of jobs -- x86-256M: ~2000 states x 220 M symbols x 2 = 330 ms against 205 ms for the 16
// serial walks side by side; BMP through the LZP coder: worse.  So one job in RCP_PILOT_EVERY is seeded and tracked first (the pilot),
// k_rcp_decide extrapolates the cost of the whole cut from what the pilot spent and what it left, compares it with the serial walk of the
// longest stream (state steps per second of the tracking kernels x seconds per symbol of the walk), and either lets the other jobs
// follow or turns every stream into one serial job.  phase: 0 = pilot jobs, 1 = the others (nothing if the decision was "serial").
#define RCP_PILOT_EVERY   16u
#define RCP_STEPS_PER_SYM 41300ull      // 2.7e12 state steps per second (tracking kernels, text-100M) x 15.3 ns per symbol (serial walk)
CR_D bool rcp_is_pilot(const RcpStream& P, uint32_t j) { return ((j - P.first_job) & (RCP_PILOT_EVERY - 1)) == RCP_PILOT_EVERY / 2; }
CR_D bool rcp_phase_skips(const RcpStream& P, uint32_t j, int phase, const RcpStats* stats) {
    if (phase == 2) return false;                                      // no pilot in this run: every job at once
    if (phase == 0) return !rcp_is_pilot(P, j);
    return rcp_is_pilot(P, j) || ((volatile const RcpStats*)stats)->decided_serial != 0;
}

// one symbol on one state: two dependent DFMA and one LOP3 (rc_dp_step, cr_rc.cuh, without the outputs)
CR_D void rcp_step(double& R, const double inv, const double f, const double nf) {
    const double T = fma(R, inv, RC_DP_MAGIC);
    const double C = fma(T, f, nf);
    R = __hiloint2double((int)rc_dp_renorm((uint32_t)__double2hiint(C)), __double2loint(C));
}
CR_D bool rcp_same(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }
CR_D double rcp_seed_state(uint32_t q, double f2) {          // 2 * norm(q * frq), exact: 2 * q * frq < 2^53
    const double C = (double)q * f2;
    return __hiloint2double((int)rc_dp_renorm((uint32_t)__double2hiint(C)), __double2loint(C));
}
struct RcpRecA { double inv, f; };
CR_D void rcp_unpack(const uint4 t, RcpRecA& a, double& nf) {
    a.inv = __hiloint2double((int)t.y, (int)t.x); a.f = __hiloint2double((int)t.w, (int)t.z); nf = -(RC_DP_TWO52 * a.f);
}
// nb symbols on K states; the records sit in shared memory (padded by one), the next one is fetched while this one is applied
template <int K> CR_D void rcp_run(double (&d)[K], const RcpRecA* __restrict__ sa, const double* __restrict__ sn, uint32_t nb) {
    RcpRecA a = sa[0]; double nf = sn[0];
#pragma unroll 2
    for (uint32_t s = 0; s < nb; s++) {
        const RcpRecA na = sa[s + 1]; const double nn = sn[s + 1];
#pragma unroll
        for (int k = 0; k < K; k++) rcp_step(d[k], a.inv, a.f, nf);
        a = na; nf = nn;
    }
}

// ---- plan: job table from the stream table.  One CTA; job counts per stream, exclusive scan, nominal starts.
__global__ void __launch_bounds__(256) k_rcp_plan(const RcStream* __restrict__ streams, uint32_t nstreams, const uint32_t* __restrict__ escord,
                                                  uint32_t job_symbols, uint32_t max_jobs, RcpStream* __restrict__ ps, RcpJob* __restrict__ jobs,
                                                  uint32_t* __restrict__ njobs_total) {
    __shared__ uint32_t wsum[8]; __shared__ uint32_t base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (uint32_t s0 = 0; s0 < nstreams; s0 += 256) {
        const uint32_t s = s0 + threadIdx.x;
        unsigned long long i0 = 0, i1 = 0; uint32_t nj = 0, is_main = 0;
        if (s < nstreams) {
            const RcStream S = streams[s];
            i0 = S.ev_begin; i1 = S.ev_end; is_main = S.is_main;
            if (S.is_main) { i0 += escord[S.ev_begin]; i1 += escord[S.ev_end]; }
            const unsigned long long n = i1 - i0;
            nj = (uint32_t)(n / job_symbols); if (nj < 2) nj = 1;
        }
        const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        uint32_t incl = nj;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
        for (uint32_t k = 0; k < 8; k++) { if (k < w) woff += wsum[k]; tot += wsum[k]; }
        const uint32_t first = base + woff + incl - nj;
        if (s < nstreams) {
            RcpStream P; P.i0 = i0; P.i1 = i1; P.first_job = first; P.is_main = is_main; P.flags = 0;
            if (first + nj > max_jobs) { nj = 0; P.flags = 1; P.first_job = 0; }      // cannot happen (host bound); the serial walk takes it
            P.njobs = nj;
            ps[s] = P;
            const unsigned long long n = i1 - i0;
            for (uint32_t c = 0; c < nj; c++) {
                RcpJob J; memset(&J, 0, sizeof J);
                J.nominal = i0 + n * c / nj; J.a = J.nominal; J.pos = J.nominal; J.end = i1; J.stream = s; J.status = c == 0 ? RCP_LIVE : 0;
                J.next = RCP_NONE; J.prev = RCP_NONE;
                jobs[first + c] = J;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *njobs_total = base < max_jobs ? base : max_jobs;
}

// ---- bounds: one warp per job c >= 1: the symbol with the largest sum in the window behind the nominal start becomes the seed.
__global__ void __launch_bounds__(128) k_rcp_bounds(const RcpStream* __restrict__ ps, RcpJob* __restrict__ jobs, const uint32_t* __restrict__ njobs_total,
                                                    const Tri* __restrict__ dense_main, const Tri* __restrict__ dense_side, RcpStats* __restrict__ stats) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= *njobs_total) return;
    const RcpJob J = jobs[j];
    const RcpStream P = ps[J.stream];
    if (P.flags || j == P.first_job) return;
    const Tri* tri = P.is_main ? dense_main : dense_side;
    // this job ends where the next nominal job starts.  The seed is the largest sum among the first RCP_WINDOW symbols; if even that
    // one leaves too many states (fresh models at the start of a chain have small sums) the search goes on through the first half
    const unsigned long long nxt = (j + 1 < P.first_job + P.njobs) ? jobs[j + 1].nominal : P.i1;
    const unsigned long long half = (nxt - J.nominal) / 2;
    const unsigned long long W = half < RCP_WINDOW ? half : RCP_WINDOW;
    uint32_t best = 0; unsigned long long at = J.nominal;
    auto scan = [&](unsigned long long lo, unsigned long long hi) {
        for (unsigned long long i = lo + lane; i < hi; i += 32) { const uint32_t s = tri[i].sum; if (s > best) { best = s; at = i; } }
        for (int d = 16; d; d >>= 1) {
            const uint32_t ob = __shfl_down_sync(0xFFFFFFFFu, best, d); const unsigned long long oa = __shfl_down_sync(0xFFFFFFFFu, at, d);
            if (ob > best || (ob == best && oa < at)) { best = ob; at = oa; }
        }
        best = __shfl_sync(0xFFFFFFFFu, best, 0); at = __shfl_sync(0xFFFFFFFFu, at, 0);
    };
    scan(J.nominal, J.nominal + W);
    if (best < (1u << 14) && half > W) scan(J.nominal + W, J.nominal + half);
    if (lane == 0) {
        const uint32_t M = best ? 0xFFFFFFFFu / best - (1u << 24) / best + 1 : 0xFFFFFFFFu;
        const bool ok = best != 0 && M <= RCP_MMAX;
        jobs[j].a = at; jobs[j].pos = at; jobs[j].status = ok ? RCP_LIVE : RCP_MERGED;
        if (ok) atomicAdd(&stats->live_jobs, 1u); else atomicAdd(&stats->merged_jobs, 1u);
    }
}

// end of job j's symbols = start of the next live job of its stream (or the stream's end).  A1 and A2 may demote jobs while
// other CTAs of the same launch call this: a stale answer only makes a job stop early, the later phases use k_rcp_links.
CR_D unsigned long long rcp_job_end(const RcpStream& P, const RcpJob* jobs, uint32_t j, uint32_t* next) {
    for (uint32_t k = j + 1; k < P.first_job + P.njobs; k++) if (((volatile const RcpJob*)jobs)[k].status == RCP_LIVE) { *next = k; return jobs[k].a; }
    *next = RCP_NONE;
    return P.i1;
}
// ---- links: once A1 and A2 are through, the set of live jobs is final
__global__ void k_rcp_links(const RcpStream* __restrict__ ps, RcpJob* __restrict__ jobs, const uint32_t* __restrict__ njobs_total) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *njobs_total) return;
    const RcpJob J = jobs[j];
    if (J.status != RCP_LIVE) return;
    const RcpStream P = ps[J.stream];
    uint32_t nx = RCP_NONE, pv = RCP_NONE;
    const unsigned long long end = rcp_job_end(P, jobs, j, &nx);
    for (uint32_t k = j; k-- > P.first_job;) if (jobs[k].status == RCP_LIVE) { pv = k; break; }
    jobs[j].end = end; jobs[j].next = nx; jobs[j].prev = pv;
}

// block-wide exclusive scan of one value per thread (THREADS <= 1024); returns the offset, *total = sum.  Two barriers.
template <int THREADS> CR_D uint32_t rcp_block_scan(uint32_t v, uint32_t* __restrict__ swarp, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
    __syncthreads();                       // swarp may still be read from the previous scan
    if (lane == 31) swarp[w] = incl;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < THREADS / 32; k++) { const uint32_t x = swarp[k]; if (k < (int)w) woff += x; tot += x; }
    *total = tot;
    return woff + incl - v;
}

// ---- A1 seed: the states possible behind the seed symbol, enumerated implicitly, S symbols each, equal neighbours dropped.
#define RCP_SEED_THREADS 1024
__global__ void __launch_bounds__(RCP_SEED_THREADS) k_rcp_seed(const RcpStream* __restrict__ ps, RcpJob* jobs, const uint32_t* __restrict__ njobs_total,
                                                               const Tri* __restrict__ dense_main, const Tri* __restrict__ dense_side,
                                                               const uint4* __restrict__ cin_main, const uint4* __restrict__ cin_side,
                                                               double* __restrict__ listA, int phase, RcpStats* __restrict__ stats) {
    __shared__ RcpRecA sa[RCP_SMAX + 1];
    __shared__ double sn[RCP_SMAX + 1];
    __shared__ uint32_t swarp[32];
    __shared__ double sfirst;
    const uint32_t j = blockIdx.x, tid = threadIdx.x;
    if (j >= *njobs_total) return;
    const RcpJob J = jobs[j];
    const RcpStream P = ps[J.stream];
    if (P.flags || J.status != RCP_LIVE || j == P.first_job || rcp_phase_skips(P, j, phase, stats)) return;
    uint32_t nextj;
    const unsigned long long end = rcp_job_end(P, jobs, j, &nextj);
    if (nextj == RCP_NONE) return;                                  // last live job of its stream: nobody needs its exit set
    const Tri seed = (P.is_main ? dense_main : dense_side)[J.a];
    const uint4* cin = P.is_main ? cin_main : cin_side;
    const uint32_t sum = seed.sum, qlo = (1u << 24) / sum, qhi = 0xFFFFFFFFu / sum, M = qhi - qlo + 1;
    const double f0 = 2.0 * (double)(seed.frq & 0x7FFFFFFFu);
    const uint32_t per = (M + RCP_SEED_THREADS - 1) / RCP_SEED_THREADS;
    const uint32_t lo = tid * per < M ? tid * per : M, hi = lo + per < M ? lo + per : M;
    double* out = listA + (size_t)j * RCP_CAP_A;
    uint32_t S = RCP_S0, staged = 0;
    unsigned long long work = 0;
    for (;;) {
        if (J.a + 1 + S > end) S = (uint32_t)(end - J.a - 1);
        for (uint32_t i = staged + tid; i < S + 1; i += RCP_SEED_THREADS) {
            RcpRecA a; double nf;
            rcp_unpack(J.a + 1 + i < end ? cin[J.a + 1 + i] : make_uint4(0, 0x3FE00000u, 0, 0x40000000u), a, nf);
            sa[i] = a; sn[i] = nf;
        }
        staged = S + 1;
        __syncthreads();
        // state of element 0 after S symbols (the seam: the last elements wrap around onto the first)
        if (tid == 0) { double r[1] = { rcp_seed_state(qlo, f0) }; rcp_run<1>(r, sa, sn, S); sfirst = r[0]; }
        __syncthreads();
        const double first = sfirst;
        uint32_t off = 0, total = 0;
        for (int pass = 0; pass < 2; pass++) {
            uint32_t cnt = 0;
            double prev = first;
            if (lo > 0 && lo < hi) { double r[1] = { rcp_seed_state(qlo + lo - 1, f0) }; rcp_run<1>(r, sa, sn, S); prev = r[0]; }   // the element in front of this slice
            for (uint32_t b = lo; b < hi; b += 8) {
                double d[8];
#pragma unroll
                for (int k = 0; k < 8; k++) d[k] = rcp_seed_state(qlo + (b + k < hi ? b + k : hi - 1), f0);
                rcp_run<8>(d, sa, sn, S);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const bool uniq = b + k < hi && (b + k == 0 || (!rcp_same(d[k], prev) && !rcp_same(d[k], first)));
                    if (uniq) { if (pass) out[off + cnt] = d[k]; cnt++; }
                    prev = d[k];
                }
            }
            if (pass == 0) {
                off = rcp_block_scan<RCP_SEED_THREADS>(cnt, swarp, &total);
                if (total > RCP_CAP_A) break;
            }
        }
        work += 2ull * (hi - lo) * S;
        if (total <= RCP_CAP_A) {
            if (tid == 0) { jobs[j].count = total; jobs[j].pos = J.a + 1 + S; }
            break;
        }
        if (S >= RCP_SMAX || J.a + 1 + S >= end) {                   // does not shrink: this job is merged into the one in front of it
            if (tid == 0) { jobs[j].status = RCP_MERGED; jobs[j].count = 0; atomicAdd(&stats->demoted_jobs, 1u); }
            break;
        }
        if (tid == 0) atomicAdd(&stats->seed_retries, 1u);
        S *= 2;
        __syncthreads();
    }
    if (work) { atomicAdd(&stats->state_steps, work); atomicAdd(&stats->phase_steps[0], work); }
}

// ---- A2 / A3 / B: explicit state lists stepped in registers (K per thread, blocked layout), merged values dropped.
//   mode 0 (A2): listA[j] (count states at pos) -> until count <= RCP_CAP_E, back into listA[j]; a job that reaches its end with
//                more states than that is merged into the job in front of it
//   mode 1 (A3): listA[j] (or the single start state for the first job of a stream) -> to the job's end, E[j]
//   mode 2 (B) : E[previous live job] stepped through this job's symbols WITHOUT merging -> F[j][i] = where candidate i ends up
enum { RCP_MODE_MID = 0, RCP_MODE_LATE = 1, RCP_MODE_FOLLOW = 2 };
template <int THREADS, int K, int BATCH, bool DEDUP>
__global__ void __launch_bounds__(THREADS) k_rcp_track(RcpStream* ps, RcpJob* jobs, const uint32_t* __restrict__ njobs_total,
                                                       const uint4* __restrict__ cin_main, const uint4* __restrict__ cin_side,
                                                       double* __restrict__ listA, double* __restrict__ E, double* __restrict__ F, int mode, int phase, RcpStats* __restrict__ stats) {
    static_assert(THREADS >= BATCH, "one record per thread and batch");
    extern __shared__ double sbuf[];                 // DEDUP: THREADS * K doubles for the compaction
    __shared__ RcpRecA sa[BATCH + 1];
    __shared__ double sn[BATCH + 1];
    __shared__ uint32_t swarp[32];
    __shared__ double slast[32];
    __shared__ double sfirst;
    const uint32_t j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (j >= *njobs_total) return;
    const RcpJob J = jobs[j];
    if (J.status != RCP_LIVE) return;
    const RcpStream P = ps[J.stream];
    if (P.flags) return;
    const bool first_job = j == P.first_job;
    const double* in; uint32_t n; unsigned long long pos, end;
    if (mode == RCP_MODE_MID) {
        if (first_job || rcp_phase_skips(P, j, phase, stats) || J.count <= RCP_CAP_E) return;
        uint32_t nextj;
        end = rcp_job_end(P, jobs, j, &nextj);
        if (nextj == RCP_NONE) return;
        in = listA + (size_t)j * RCP_CAP_A; n = J.count; pos = J.pos;
    } else {
        if (J.next == RCP_NONE) return;              // last live job of the stream: nobody needs its exit set
        end = J.end;
        if (mode == RCP_MODE_FOLLOW) {
            if (first_job) return;
            in = E + (size_t)J.prev * RCP_CAP_E; n = jobs[J.prev].ecount; pos = J.a;
        } else if (first_job) return;                 // k_rcp_emit(RCP_EMIT_FIRST) has written its one-element exit set
        else { in = listA + (size_t)j * RCP_CAP_A; n = J.count; pos = J.pos; }
    }
    if (n == 0 || n > (uint32_t)THREADS * K || pos > end) { if (tid == 0) atomicOr(&ps[J.stream].flags, 8u); return; }
    const uint4* cin = P.is_main ? cin_main : cin_side;
    double d[K];
#pragma unroll
    for (int k = 0; k < K; k++) { const uint32_t i = tid * K + k; d[k] = in ? in[i < n ? i : n - 1] : RC_DP_R0; }
    unsigned long long work = 0;
    uint32_t batches = 0;
    const uint4 idle = make_uint4(0, 0x3FE00000u, 0, 0x40000000u);
    uint4 nxt = idle;
    if (tid <= BATCH && pos + tid < end) nxt = cin[pos + tid];
    while (pos < end) {
        __syncthreads();                                               // the previous batch has been consumed
        if (tid <= BATCH) { RcpRecA a; double nf; rcp_unpack(nxt, a, nf); sa[tid] = a; sn[tid] = nf; }
        __syncthreads();
        const uint32_t nb = end - pos < BATCH ? (uint32_t)(end - pos) : BATCH;
        nxt = idle;
        if (tid <= BATCH && pos + BATCH + tid < end) nxt = cin[pos + BATCH + tid];
        if (tid * K < n) {
            rcp_run<K>(d, sa, sn, nb);
            if (lane == 0) work += (unsigned long long)nb * (n - tid * K < 32u * K ? n - tid * K : 32u * K);
        }
        pos += nb;
        batches++;
        // merged states are dropped after every batch while the list still has to shrink (MID), otherwise after every fourth batch and
        // at the end: the compaction (a block-wide scan and a trip through shared memory) costs as much as a batch of 128 symbols, and
        // a set that is already small loses elements slowly (~1 / sqrt(symbols))
        if (DEDUP && (mode == RCP_MODE_MID || (batches & 3u) == 0 || pos >= end)) {
            // drop every element equal to its predecessor (or, wrapping around, to element 0); order is kept
            if (lane == 31) slast[w] = d[K - 1];
            if (tid == 0) sfirst = d[0];
            __syncthreads();
            const double first = sfirst;
            double prev = __shfl_up_sync(0xFFFFFFFFu, d[K - 1], 1);
            if (lane == 0) prev = w ? slast[w - 1] : first;
            uint32_t keep = 0, cnt = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t i = tid * K + k;
                const bool uniq = i < n && (i == 0 || (!rcp_same(d[k], prev) && !rcp_same(d[k], first)));
                keep |= (uint32_t)uniq << k; cnt += uniq;
                prev = d[k];
            }
            uint32_t total;
            uint32_t off = rcp_block_scan<THREADS>(cnt, swarp, &total);
            if (total < n) {
#pragma unroll
                for (int k = 0; k < K; k++) if (keep >> k & 1) sbuf[off++] = d[k];
                __syncthreads();
                n = total;
#pragma unroll
                for (int k = 0; k < K; k++) { const uint32_t i = tid * K + k; d[k] = sbuf[i < n ? i : n - 1]; }
            }
            if (mode == RCP_MODE_MID && n <= RCP_CAP_E) break;
        }
    }
    if (mode == RCP_MODE_MID && n > RCP_CAP_E) {                        // did not shrink far enough within the job: merge it
        if (tid == 0) { jobs[j].status = RCP_MERGED; atomicAdd(&stats->demoted_jobs, 1u); }
        return;
    }
    double* out = mode == RCP_MODE_MID ? listA + (size_t)j * RCP_CAP_A : mode == RCP_MODE_LATE ? E + (size_t)j * RCP_CAP_E : F + (size_t)j * RCP_CAP_E;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) { const uint32_t i = tid * K + k; if (i < n) out[i] = d[k]; }
    if (tid == 0) {
        if (mode == RCP_MODE_MID) { jobs[j].count = n; jobs[j].pos = pos; }
        else if (mode == RCP_MODE_LATE) { jobs[j].ecount = n; atomicMax(&stats->max_e, n); }
    }
    if (lane == 0 && work) { atomicAdd(&stats->state_steps, work); atomicAdd(&stats->phase_steps[1 + mode], work); }
}

// ---- C resolve: one warp per stream walks its live jobs: entry of the next job = F[this job][index of this job's entry in E[previous]]
__global__ void __launch_bounds__(32) k_rcp_resolve(RcpStream* ps, uint32_t nstreams, RcpJob* jobs, const double* __restrict__ E, const double* __restrict__ F) {
    const uint32_t s = blockIdx.x, lane = threadIdx.x;
    if (s >= nstreams) return;
    const RcpStream P = ps[s];
    if (P.flags || P.njobs == 0) return;
    uint32_t cur = P.first_job, prev = RCP_NONE;
    double entry = RC_DP_R0;
    if (lane == 0) jobs[cur].entry = entry;
    for (;;) {
        const uint32_t nxt = jobs[cur].next;
        if (nxt == RCP_NONE) break;
        double ne = 0.0; bool ok = false;
        if (prev == RCP_NONE) {                        // first job: its exit set is the one state the known start leads to
            ok = jobs[cur].ecount == 1; ne = E[(size_t)cur * RCP_CAP_E];
        } else {
            const uint32_t n = jobs[prev].ecount;
            const double* e = E + (size_t)prev * RCP_CAP_E;
            uint32_t found = RCP_NONE;
            for (uint32_t i0 = 0; i0 < n && found == RCP_NONE; i0 += 128) {
                bool hit[4];
#pragma unroll
                for (int u = 0; u < 4; u++) hit[u] = i0 + 32 * u + lane < n && rcp_same(e[i0 + 32 * u + lane], entry);
#pragma unroll
                for (int u = 0; u < 4; u++) { const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit[u]); if (m && found == RCP_NONE) found = i0 + 32 * u + __ffs(m) - 1; }
            }
            if (found != RCP_NONE) { ok = true; ne = F[(size_t)cur * RCP_CAP_E + found]; }
        }
        if (!ok) { if (lane == 0) atomicOr(&ps[s].flags, 32u); return; }
        entry = ne; prev = cur; cur = nxt;
        if (lane == 0) jobs[cur].entry = entry;
    }
}

// ---- D emit: the serial walk per job from its true entry state; checks the link to the next job.
//   RCP_EMIT_JOBS   every live job but the first of its stream
//   RCP_EMIT_WHOLE  the fallback: one warp per FLAGGED stream walks the whole stream from the coder's initial range
//   RCP_EMIT_FIRST  the first job of every stream (its entry is the coder's initial range); its exit state becomes E[job] (one element),
//                   so no tracking pass has to walk these symbols a second time
// Lane 0 runs the chain and nothing else: per symbol two dependent DFMA, one LOP3 on the high word and one 8-byte store of T (whose low
// word is the quotient).  Records come from shared memory four symbols ahead; -(2^52 * 2 frq) is computed by the staging lanes; the
// top-bit index is recomputed from T by all lanes afterwards (C = fma(T, 2 frq, nf) again, off the chain).  A full batch is straight-line
// code.  Measured: profiles/round2_summary.md section 1.
enum { RCP_EMIT_JOBS = 0, RCP_EMIT_WHOLE = 1, RCP_EMIT_FIRST = 2 };
CR_D void rcp_chain_step(double& R, const uint4 t, const double nf, double& T) {
    T = fma(R, rc_dp_join(t.y, t.x), RC_DP_MAGIC);
    const double C = fma(T, rc_dp_join(t.w, t.z), nf);
    R = __hiloint2double((int)rc_dp_renorm((uint32_t)__double2hiint(C)), __double2loint(C));
}
__global__ void __launch_bounds__(128) k_rcp_emit(RcpStream* ps, uint32_t nstreams, RcpJob* jobs, const uint32_t* __restrict__ njobs_total,
                                                  const uint4* __restrict__ cin_main, const uint4* __restrict__ cin_side,
                                                  uint32_t* __restrict__ q_main, uint32_t* __restrict__ sh_main, uint32_t* __restrict__ q_side, uint32_t* __restrict__ sh_side,
                                                  int which, double* __restrict__ E, RcpStats* __restrict__ stats) {
    __shared__ uint4 stage[4][RC_BATCH + 8];
    __shared__ double snf[4][RC_BATCH + 8];
    __shared__ double sT[4][RC_BATCH];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long i0, i1; double R; uint32_t sidx, nextj = RCP_NONE; bool is_main;
    if (which == RCP_EMIT_WHOLE) {
        if (g >= nstreams) return;
        const RcpStream P = ps[g];
        if (!P.flags) return;
        if (lane == 0) atomicAdd(&stats->flagged_streams, 1u);
        i0 = P.i0; i1 = P.i1; R = RC_DP_R0; sidx = g; is_main = P.is_main != 0;
    } else {
        if (g >= *njobs_total) return;
        const RcpJob J = jobs[g];
        if (J.status != RCP_LIVE) return;
        const RcpStream P = ps[J.stream];
        if (P.flags) return;
        const bool first = g == P.first_job;
        if (first != (which == RCP_EMIT_FIRST)) return;
        i1 = J.end; nextj = J.next;
        i0 = first ? P.i0 : J.a; R = first ? RC_DP_R0 : J.entry; sidx = J.stream; is_main = P.is_main != 0;
    }
    const uint4* tri = is_main ? cin_main : cin_side;
    uint32_t* qo = is_main ? q_main : q_side;
    uint32_t* so = is_main ? sh_main : sh_side;
    uint4 r0 = make_uint4(1, 1, 1, 1), r1 = r0;
    if (i0 + lane < i1) r0 = tri[i0 + lane];
    if (i0 + 32 + lane < i1) r1 = tri[i0 + 32 + lane];
    if (lane < 8) { stage[w][RC_BATCH + lane] = make_uint4(0, 0x3FE00000u, 0, 0x40000000u); snf[w][RC_BATCH + lane] = 0.0; }    // read ahead only, never applied
    for (unsigned long long base = i0; base < i1; base += RC_BATCH) {
        stage[w][lane] = r0; stage[w][lane + 32] = r1;
        const double nfa = -(RC_DP_TWO52 * rc_dp_join(r0.w, r0.z)), nfb = -(RC_DP_TWO52 * rc_dp_join(r1.w, r1.z));
        const double fa = rc_dp_join(r0.w, r0.z), fb = rc_dp_join(r1.w, r1.z);
        snf[w][lane] = nfa; snf[w][lane + 32] = nfb;
        __syncwarp();
        const unsigned long long nb = base + RC_BATCH;
        if (nb + lane < i1) r0 = tri[nb + lane];
        if (nb + 32 + lane < i1) r1 = tri[nb + 32 + lane];
        const uint32_t cnt = i1 - base < RC_BATCH ? (uint32_t)(i1 - base) : RC_BATCH;
        if (lane == 0) {
            const uint4* st = stage[w]; const double* sn = snf[w]; double* so_t = sT[w];
            if (cnt == RC_BATCH) {
                uint4 c0 = st[0], c1 = st[1], c2 = st[2], c3 = st[3];
                double n0 = sn[0], n1 = sn[1], n2 = sn[2], n3 = sn[3];
#pragma unroll
                for (uint32_t j = 0; j < RC_BATCH; j += 4) {
                    const uint4 d0 = st[j + 4], d1 = st[j + 5], d2 = st[j + 6], d3 = st[j + 7];
                    const double m0 = sn[j + 4], m1 = sn[j + 5], m2 = sn[j + 6], m3 = sn[j + 7];
                    double T0, T1, T2, T3;
                    rcp_chain_step(R, c0, n0, T0); so_t[j] = T0;
                    rcp_chain_step(R, c1, n1, T1); so_t[j + 1] = T1;
                    rcp_chain_step(R, c2, n2, T2); so_t[j + 2] = T2;
                    rcp_chain_step(R, c3, n3, T3); so_t[j + 3] = T3;
                    c0 = d0; c1 = d1; c2 = d2; c3 = d3; n0 = m0; n1 = m1; n2 = m2; n3 = m3;
                }
            } else {
                for (uint32_t j = 0; j < cnt; j++) { double T; rcp_chain_step(R, st[j], sn[j], T); so_t[j] = T; }
            }
        }
        __syncwarp();
        {   // quotient = low word of T; top-bit index of q * frq = exponent of C
            const double Ta = sT[w][lane], Tb = sT[w][lane + 32];
            const double Ca = fma(Ta, fa, nfa), Cb = fma(Tb, fb, nfb);
            if (lane < cnt) { qo[base + lane] = (uint32_t)__double2loint(Ta); so[base + lane] = ((uint32_t)__double2hiint(Ca) >> 20) - 1024u; }
            if (lane + 32 < cnt) { qo[base + lane + 32] = (uint32_t)__double2loint(Tb); so[base + lane + 32] = ((uint32_t)__double2hiint(Cb) >> 20) - 1024u; }
        }
        __syncwarp();
    }
    if (lane != 0) return;
    if (which == RCP_EMIT_FIRST) { if (nextj != RCP_NONE) { E[(size_t)g * RCP_CAP_E] = R; jobs[g].ecount = 1; } }
    else if (which == RCP_EMIT_JOBS && nextj != RCP_NONE && !rcp_same(R, jobs[nextj].entry)) atomicOr(&ps[sidx].flags, 64u);   // the jobs do not link up
}

// ---- decide (one CTA): see RCP_PILOT_EVERY above
__global__ void __launch_bounds__(256) k_rcp_decide(const RcpStream* __restrict__ ps, uint32_t nstreams, RcpJob* jobs, const uint32_t* __restrict__ njobs_total, uint32_t job_symbols, RcpStats* __restrict__ stats) {
    __shared__ unsigned long long s_sum, s_total, s_longest; __shared__ uint32_t s_pilots, s_alive, s_serial;
    if (threadIdx.x == 0) { s_sum = 0; s_total = 0; s_longest = 0; s_pilots = 0; s_alive = 0; s_serial = 0; }
    __syncthreads();
    const uint32_t nj = *njobs_total;
    for (uint32_t j = threadIdx.x; j < nj; j += 256) {
        const RcpJob J = jobs[j];
        const RcpStream P = ps[J.stream];
        if (P.flags || j == P.first_job || !rcp_is_pilot(P, j)) continue;
        atomicAdd(&s_pilots, 1u);
        // the average size of the set the tracking passes will carry through this job.  The set shrinks roughly as 1 / sqrt(symbols
        // behind the seed) (profiles/round2_summary.md section 0): it had `count` states `t` symbols behind the seed, so over a job of
        // T symbols it averages count * 2 sqrt(t / T) (at most count).  A job without a usable seed counts as a full set.
        uint32_t left = RCP_CAP_E;
        if (J.status == RCP_LIVE && J.count > 0) {
            atomicAdd(&s_alive, 1u);
            const uint32_t c = J.count < RCP_CAP_E ? J.count : RCP_CAP_E;
            const float f = 2.0f * sqrtf((float)(J.pos - J.a + 1) / (float)job_symbols);
            left = f < 1.0f ? (uint32_t)((float)c * f) + 1u : c;
        }
        atomicAdd(&s_sum, (unsigned long long)left);
    }
    for (uint32_t s = threadIdx.x; s < nstreams; s += 256) {
        const RcpStream P = ps[s];
        if (P.flags) continue;
        atomicAdd(&s_total, P.i1 - P.i0);
        atomicMax(&s_longest, P.i1 - P.i0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool serial = false;
        unsigned long long est = 0, ser = s_longest * RCP_STEPS_PER_SYM;
        if (s_pilots >= 4) {                                           // (fewer: small windows, nothing to lose either way)
            const unsigned long long spent = stats->phase_steps[0] + stats->phase_steps[1];
            est = spent * RCP_PILOT_EVERY + 2ull * (s_sum / s_pilots) * s_total;
            serial = est > ser || s_alive * 2 < s_pilots;
        }
        stats->decided_serial = serial ? 1u : 0u; stats->pilot_jobs = s_pilots; stats->est_steps = est; stats->serial_equiv = ser;
        s_serial = serial ? 1u : 0u;
    }
    __syncthreads();
    if (!s_serial) return;
    for (uint32_t j = threadIdx.x; j < nj; j += 256) {                 // one serial job per stream
        const RcpStream P = ps[jobs[j].stream];
        if (!P.flags && j != P.first_job && jobs[j].status == RCP_LIVE) jobs[j].status = RCP_MERGED;
    }
}

// ---- prune: a stream whose live jobs leave a span longer than a third of the stream gains little from the cut and pays for the
// tracking of the jobs around it (BMP data: sums below 2^14 almost everywhere, a handful of seeds survive): it becomes one serial job.
__global__ void k_rcp_prune(const RcpStream* __restrict__ ps, uint32_t nstreams, RcpJob* jobs, int force_serial, RcpStats* __restrict__ stats) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const RcpStream P = ps[s];
    if (P.flags || P.njobs < 2) return;
    unsigned long long last = P.i0, longest = 0; uint32_t live = 0;
    for (uint32_t k = P.first_job + 1; k < P.first_job + P.njobs; k++) if (jobs[k].status == RCP_LIVE) {
        if (jobs[k].a - last > longest) longest = jobs[k].a - last;
        last = jobs[k].a; live++;
    }
    if (P.i1 - last > longest) longest = P.i1 - last;
    if (!live || (!force_serial && longest * 3 <= P.i1 - P.i0)) return;
    for (uint32_t k = P.first_job + 1; k < P.first_job + P.njobs; k++) if (jobs[k].status == RCP_LIVE) { jobs[k].status = RCP_MERGED; atomicAdd(&stats->demoted_jobs, 1u); }
}

// host side: all buffers of the parallel chain
struct RcPar {
    DevBuf b_ps, b_jobs, b_njobs, b_listA, b_E, b_F, b_stats;
    uint32_t job_symbols = 0;        // symbols per job; 0 = chosen from the window's size
    int late_cfg = 0;                // tuning: thread / register layout of the A3 and B kernels
    int serial_only = -1;            // 1: one serial job per stream; 0: always cut; -1: cut unless `crowded` (several handles share the device)
    bool crowded = false;            // set by crgpu_compress_batch: the serial walks of different handles overlap, the tracking kernels do not
    bool attr_done = false;
    RcpStats last = {};
    void release() { DevBuf* all[] = { &b_ps, &b_jobs, &b_njobs, &b_listA, &b_E, &b_F, &b_stats }; for (DevBuf* b : all) b->release(); }
    // q / top-bit index of every symbol of every stream, as k_range_chain<7> writes them
    int run(cudaStream_t stream, const RcStream* d_streams, uint32_t nstreams, const uint32_t* d_escord, uint64_t ntm, uint64_t nts,
            const Tri* dense_main, const Tri* dense_side, const uint4* cin_main, const uint4* cin_side,
            uint32_t* q_main, uint32_t* sh_main, uint32_t* q_side, uint32_t* sh_side, bool want_stats) {
        if (nstreams == 0) return CRGPU_OK;
        const bool serial = serial_only == 1 || (serial_only < 0 && crowded);
        uint32_t T = job_symbols;
        if (T == 0) {                // about eight jobs per SM when the window is large; a job of fewer than 16384 symbols does not pay for its seed
            const uint64_t t = (ntm + nts) / 1184;
            T = t < 16384 ? 16384u : t > 65536 ? 65536u : (uint32_t)((t + 4095) & ~4095ull);
        }
        if (T < 4096) T = 4096;
        if (serial) T = 0xFFFFFFFFu;                                   // one job per stream
        const uint64_t maxjobs64 = nstreams + (serial ? 0 : (ntm + nts) / T) + 1;
        if (maxjobs64 > (1u << 20)) return CRGPU_ERR_UNSUPPORTED;
        const uint32_t maxjobs = (uint32_t)maxjobs64;
        CR_TRY(b_ps.reserve((size_t)nstreams * sizeof(RcpStream))); CR_TRY(b_jobs.reserve((size_t)maxjobs * sizeof(RcpJob)));
        CR_TRY(b_njobs.reserve(16)); CR_TRY(b_stats.reserve(sizeof(RcpStats)));
        if (!serial) { CR_TRY(b_listA.reserve((size_t)maxjobs * RCP_CAP_A * 8)); CR_TRY(b_F.reserve((size_t)maxjobs * RCP_CAP_E * 8)); }
        CR_TRY(b_E.reserve((size_t)maxjobs * RCP_CAP_E * 8));
        CR_CUDA(cudaMemsetAsync(b_stats.p, 0, sizeof(RcpStats), stream));
        RcpStream* ps = b_ps.as<RcpStream>(); RcpJob* jobs = b_jobs.as<RcpJob>(); uint32_t* nj = b_njobs.as<uint32_t>(); RcpStats* st = b_stats.as<RcpStats>();
        double* LA = b_listA.as<double>(); double* E = b_E.as<double>(); double* F = b_F.as<double>();
        CR_LAUNCH(k_rcp_plan, dim3(1), dim3(256), stream, d_streams, nstreams, d_escord, T, maxjobs, ps, jobs, nj);
        if (!serial) {
        CR_LAUNCH(k_rcp_bounds, dim3(cr_div_up((size_t)maxjobs * 32, 128)), dim3(128), stream, ps, jobs, nj, dense_main, dense_side, st);
        if (!attr_done) {                            // per handle: a handle is bound to one device
            CR_CUDA(cudaFuncSetAttribute(k_rcp_track<512, 32, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 32 * 8));
            attr_done = true;
        }
#define RCP_TRACK(TH, KK, BB, DD, SMEM, MODE, PHASE)                                                                                \
        do { __atomic_fetch_add(&g_cr_launches, 1ull, __ATOMIC_RELAXED);                                                             \
             k_rcp_track<TH, KK, BB, DD><<<dim3(maxjobs), dim3(TH), (SMEM), stream>>>(ps, jobs, nj, cin_main, cin_side, LA, E, F, MODE, PHASE, st); \
             CR_CUDA(cudaGetLastError()); } while (0)
        const bool pilot = serial_only < 0 && maxjobs >= 8 * RCP_PILOT_EVERY;       // "always cut" (rc_serial 0) and small windows: no pilot
        if (pilot) {
            CR_LAUNCH(k_rcp_seed, dim3(maxjobs), dim3(RCP_SEED_THREADS), stream, ps, jobs, nj, dense_main, dense_side, cin_main, cin_side, LA, 0, st);
            RCP_TRACK(512, 32, 16, true, 512 * 32 * 8, RCP_MODE_MID, 0);
            CR_LAUNCH(k_rcp_decide, dim3(1), dim3(256), stream, ps, nstreams, jobs, nj, T, st);
            CR_LAUNCH(k_rcp_seed, dim3(maxjobs), dim3(RCP_SEED_THREADS), stream, ps, jobs, nj, dense_main, dense_side, cin_main, cin_side, LA, 1, st);
            RCP_TRACK(512, 32, 16, true, 512 * 32 * 8, RCP_MODE_MID, 1);
        } else {
            CR_LAUNCH(k_rcp_seed, dim3(maxjobs), dim3(RCP_SEED_THREADS), stream, ps, jobs, nj, dense_main, dense_side, cin_main, cin_side, LA, 2, st);
            RCP_TRACK(512, 32, 16, true, 512 * 32 * 8, RCP_MODE_MID, 2);
        }
        CR_LAUNCH(k_rcp_prune, dim3(cr_div_up(nstreams, 64)), dim3(64), stream, ps, nstreams, jobs, 0, st);
        }
        CR_LAUNCH(k_rcp_links, dim3(cr_div_up(maxjobs, 128)), dim3(128), stream, ps, jobs, nj);
        const dim3 gemit(cr_div_up((size_t)maxjobs * 32, 128));
        // first jobs: their entry is known, the walk that emits them also yields their exit state
        CR_LAUNCH(k_rcp_emit, gemit, dim3(128), stream, ps, nstreams, jobs, nj, cin_main, cin_side, q_main, sh_main, q_side, sh_side, (int)RCP_EMIT_FIRST, E, st);
        if (!serial) {
        if (late_cfg == 1) { RCP_TRACK(256, 8, 128, true, 2048 * 8, RCP_MODE_LATE, 2); RCP_TRACK(256, 8, 128, false, 0, RCP_MODE_FOLLOW, 2); }
        else if (late_cfg == 2) { RCP_TRACK(128, 16, 128, true, 2048 * 8, RCP_MODE_LATE, 2); RCP_TRACK(128, 16, 128, false, 0, RCP_MODE_FOLLOW, 2); }
        else if (late_cfg == 3) { RCP_TRACK(256, 8, 256, true, 2048 * 8, RCP_MODE_LATE, 2); RCP_TRACK(256, 8, 256, false, 0, RCP_MODE_FOLLOW, 2); }
        else if (late_cfg == 4) { RCP_TRACK(512, 4, 128, true, 2048 * 8, RCP_MODE_LATE, 2); RCP_TRACK(512, 4, 128, false, 0, RCP_MODE_FOLLOW, 2); }
        else { RCP_TRACK(256, 8, 256, true, 2048 * 8, RCP_MODE_LATE, 2); RCP_TRACK(256, 8, 256, false, 0, RCP_MODE_FOLLOW, 2); }
        CR_LAUNCH(k_rcp_resolve, dim3(nstreams), dim3(32), stream, ps, nstreams, jobs, E, F);
        CR_LAUNCH(k_rcp_emit, gemit, dim3(128), stream, ps, nstreams, jobs, nj, cin_main, cin_side, q_main, sh_main, q_side, sh_side, (int)RCP_EMIT_JOBS, E, st);
        }
#undef RCP_TRACK
        CR_LAUNCH(k_rcp_emit, dim3(cr_div_up((size_t)nstreams * 32, 128)), dim3(128), stream, ps, nstreams, jobs, nj, cin_main, cin_side, q_main, sh_main, q_side, sh_side, (int)RCP_EMIT_WHOLE, E, st);
        if (want_stats) {
            CR_CUDA(cudaMemcpyAsync(&last, st, sizeof(RcpStats), cudaMemcpyDeviceToHost, stream));
            CR_CUDA(cudaStreamSynchronize(stream));
        }
        return CRGPU_OK;
    }
};
#endif
