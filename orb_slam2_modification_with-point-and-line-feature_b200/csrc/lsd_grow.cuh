// lsd_grow.cuh — k_lsd_grow: the ordered main loop of LSD (flsd() of imgproc/src/lsd.cpp, OpenCV 4.13 behaviour) on sm_100a:
// region_grow(), region2rect(), refine(), reduce_region_radius().  Included by line_kernels.cu.
//
// Region growing is ordered (seeds by gradient bin, one shared USED map, a running region angle), but regions that do not
// touch the same pixels commute.  A persistent CTA takes one frame at a time from a global counter and runs the ordered loop
// as a window of speculative transactions with in-order commit.  ONE THREAD PER REGION: nothing in the life of a region is
// warp-cooperative, so the 32 lanes of a warp work on 32 different regions, a frame keeps a few hundred regions in flight on
// one SM, and every floating-point sum is formed in the reference's order simply because a thread runs the reference's loop.
//
//   issuer (1 warp)      walks the seed list 32 seeds at a time and gives every seed that is unused in the committed map the
//                        next ticket (a slot of the window in shared memory).
//   grower threads       take the next ticket and run the region's whole life:
//                          region_grow: pop the next region point, load the eight 16-byte pixel records of its neighbours
//                            together (one round trip to L2 per point), accept the aligned ones in row-major order, each
//                            acceptance updating the float sums and the region angle (fastAtan2);
//                          region2rect / refine (angle statistics, re-growth with the new tolerance, reduce_region_radius) as
//                            plain sequential loops over the region list.
//                        The pixels a thread holds are the ones that carry its stamp (attempt << 24 | ticket + 1, written by
//                        compare-and-swap from the value the thread saw).  The last 32 points live in the ticket's slot in shared
//                        memory; from min_reg_size points on the list also goes to a buffer in global memory.
//   committer (1 warp)   strictly in ticket order.  Runs of small regions (no rectangle: the points are in the slot) are
//                        committed up to 32 tickets at a time, one per lane; the first ticket of a run that needs more (a region
//                        with a buffer, a failed speculation) ends the batch and is handled alone.
//
// What makes a speculative growth the sequential one (checked at commit, when every earlier ticket is committed):
//   (a) no pixel the growth ever accepted is committed by then (the committed map only holds pixels of earlier tickets), and
//   (b) every pixel it skipped BECAUSE an earlier, then uncommitted ticket held it did end up in that ticket's region: the
//       growth records the stamps it deferred to (at most four different ones), and the committer checks that each of those
//       tickets committed exactly the speculation that wrote the stamp (fring);
//   and the region list has no duplicates: a thread treats "stamp == mine" as "already in my region", so a ticket that loses a
//   pixel to an earlier ticket is told (its poison word) and does not trust its growth.
// A ticket that fails any of this (or gave up: too many dependencies, a lost compare-and-swap, no buffer) is grown again by
// the committer at its turn, when everything before it is committed: that growth IS the sequential one.  Committed pixels are
// never released, tickets follow the seed order, so the result is the reference's whatever the stamps say.
// rect_improve only reads the angle map: k_lsd_nfa, afterwards.
#pragma once

namespace pl {

struct LsdQueueItem { LsdRect rec; };

constexpr int kRing = 32;        // region points a ticket slot keeps in shared memory (>= min_reg_size of any supported image)
constexpr int kPoolBufs = 512;   // region buffers per CTA (2 x kSpecCap words each: first growth | refine's re-growth); a finished region keeps
                                 // its buffer until it is committed, so the pool has to match the window of uncommitted tickets
constexpr int kSpecCap = 4096;   // capacity of a buffer half; larger regions are grown by the committer with frame-sized buffers
constexpr int kMaxDeps = 4;
constexpr int kMaxSlots = 2048;
constexpr unsigned kTicketMask = 0xffffffu;  // stamp = attempt << 24 | ticket + 1  (0 = never stamped)

enum { kSlotFree = 0, kSlotReady = 1, kSlotGrowing = 2, kSlotDone = 4 };
enum { kStDeferred = -2, kStCapacity = -1, kStNoRect = 0, kStRect = 1 };
__device__ __forceinline__ int slot_pack(int state, int status, int buf) { return state | ((status + 2) << 8) | ((buf + 1) << 16); }

// A ticket slot, in shared memory (60 bytes):
//   state[W]  state | (status + 2) << 8 | (buffer + 1) << 16
//   fin[W]    the stamp the finished region carries (0 while it is not final or when it failed)
//   poison[W] bit a: an earlier ticket took a pixel that carried this ticket's stamp of attempt a
//   hdr[W]    int4 {seed pixel | adjacency hint << 31, region size, size of the first growth when refine grew the region again (0: it
//             did not), size of that re-growth before reduce_region_radius removed points (they stay in the log)}
//   dep[W]    uint4: the stamps of earlier tickets this growth deferred to (0 = empty)
//   pts4[W]   uint4: the first four points of a region without a buffer; the others are in global memory (TicketRec)
struct TicketRec {
    unsigned int pts[kRing];
};
struct GrowLayout {
    int W;           // ticket slots, a power of two
    int window;      // tickets that may be uncommitted at once (<= W)
    int bits_words;  // words of a W*H bitmap
    int stall;       // iterations a thread waits for an earlier ticket to become final before it assumes the stamp stays
    int restarts;    // speculative attempts a ticket gets after it lost a pixel to an earlier ticket (or a compare-and-swap)
    int poll_ns;
    int debug;       // test hook (PLSLAM_LSD_DEBUG): bits 4 / 7 shadow growth of every / every large committed region, 5 no batched commit,
                     // 8 commit log, 9 shadow growth after a re-growth at the commit head
    int rings;       // point rings: one per grower thread + one for the committer
    __host__ __device__ static size_t a16(size_t v) { return (v + 15) & ~(size_t)15; }
    __host__ __device__ size_t off_hdr() const { return 0; }
    __host__ __device__ size_t off_dep() const { return (size_t)W * 16; }
    __host__ __device__ size_t off_pts4() const { return (size_t)W * 32; }
    __host__ __device__ size_t off_state() const { return (size_t)W * 48; }
    __host__ __device__ size_t off_fin() const { return off_state() + (size_t)W * 4; }
    __host__ __device__ size_t off_poison() const { return off_fin() + (size_t)W * 4; }
    __host__ __device__ size_t off_ring() const { return off_poison() + (size_t)W * 4; }
    __host__ __device__ size_t off_used() const { return off_ring() + (size_t)rings * kRing * 4; }
    __host__ __device__ size_t total() const { return off_used() + a16((size_t)bits_words * 4 + 4); }
};

struct GrowCtl {
    int ticket_next;  // tickets issued
    int grow_next;    // tickets handed to a grower
    int commit_head;  // tickets committed
    int all_issued;   // the seed list is exhausted
    int done;         // the frame is finished
    int abort_;       // watchdog
    int active;       // tickets held by grower threads
    unsigned long long free_mask[kPoolBufs / 64];
    unsigned long long stat[8];  // committed, void, regrown, committer asleep, grower-thread cycles, fit cycles, commit cycles, regrow cycles
    unsigned int why[16];        // profiling: 0 seed swallowed at take, 1 seed held by an earlier ticket, 2 too many dependencies, 3 CAS lost,
                                 // 4 poisoned, 5 capacity, 6 stalls, 7 conflict at commit, 8 dependency failed at commit, 9 dependencies taken
    long long t_start;
};

struct GrowBufs {
    const float* angdeg;
    const int* g2;
    LsdPix* rec;
    const float2* cs0;
    const unsigned int* seeds;
    const int* n_seeds;
    unsigned int* big_reg;       // [cta][2 * plane]: the committer's own growth (first growth | refine's re-growth)
    TicketRec* trec;             // [cta][kMaxSlots]
    unsigned int* gfinal;        // [cta][2 * kMaxSlots] what each recently committed ticket committed: its final stamp, or 0
    unsigned int* pool_reg;      // [cta][kPoolBufs][2][kSpecCap]
    LsdRect* pool_rect;          // [cta][kPoolBufs]
    LsdQueueItem* queue;         // [frame][seg_cap]
    int* n_rects;
    int* flags;
    long long* phase_cycles;
    int* frame_counter;
    size_t plane;
    long long watchdog_cycles;
    unsigned int* dbg_bits;   // [cta][bits_words] test hook (debug bit 4): private bitmap of the shadow growth, all-zero between uses
    int* dbg_out;             // [frame][16] first speculative region that differs from the sequential one
    int4* dbg_log;            // [frame][kDbgLogCap] test hook (debug bit 8): the committed regions in order
    int* dbg_log_n;           // [frame]
    unsigned int* dbg_scratch;  // [cta][plane]
};
constexpr int kDbgLogCap = 32768;

__device__ __forceinline__ int pool_pop_thread(unsigned long long* masks) {
#pragma unroll 1
    for (int w = 0; w < kPoolBufs / 64; w++) {
        while (true) {
            const unsigned long long m = *(volatile unsigned long long*)(masks + w);
            if (!m) break;
            const int c = __ffsll((long long)m) - 1;
            if (atomicCAS(masks + w, m, m & ~(1ull << c)) == m) return w * 64 + c;
        }
    }
    return -1;
}
__device__ __forceinline__ void pool_push(unsigned long long* masks, int b) { atomicOr(masks + (b >> 6), 1ull << (b & 63)); }

// everything a region's thread needs of its frame and CTA
struct TEnv {
    GrowCtl* ctl;
    int4* hdr;             // [W] shared memory
    uint4* dep;            // [W]
    uint4* pts4;           // [W]
    unsigned int* state;   // [W]
    unsigned int* fin;     // [W]
    unsigned int* poison;  // [W]
    unsigned int* ring;    // [rings][kRing]
    unsigned int* used;
    TicketRec* trec;       // [W] global memory
    unsigned int* gfinal;  // [2W] global memory
    int wm;
    int W, H, min_reg, stall, restarts, debug;
    LsdPix* rec;
    const float2* cs0;
    const float* ang;
    const int* g2;
    unsigned int* pool_reg;
    LsdRect* pool_rect;
    int* flag;   // the frame's error word
    bool prof;
};
// consistency checks of the grower (a violation aborts the frame with PL_ERR_CAPACITY flags = 8 | code << 4)
#define PL_LSD_CHECK(cond, code)                                                     \
    do {                                                                             \
        if (!(cond)) {                                                               \
            atomicOr(E.flag, 8 | ((code) << 4));                                     \
            *(volatile int*)&E.ctl->abort_ = 1;                                      \
        }                                                                            \
    } while (0)

// the state of a thread's region
struct TState {
    int ticket;          // -1: idle
    int phase;           // 1: waiting before it looks at its seed, 2: growing
    int n, i;            // region size, next region point to expand
    int n0;              // size of the first growth once refine grows the region again
    int buf;             // pool buffer (-1: none yet; -2: the committer's frame-sized buffers)
    int bad;             // 1: lost a pixel / a compare-and-swap (may start again), 2: capacity, 3: left to the committer
    int attempt;         // even: a first growth, odd: refine's re-growth of it (0, 1; 2, 3 after a restart ...; 62, 63 at the commit head)
    int restarts_left;
    int kstart;          // neighbour to resume the current point from
    int pix;             // the seed
    int wait_u;          // >= 0: waiting for this earlier ticket to become final
    bool wait_hard;      // ... until it is committed, however long it takes (the ticket at the commit head)
    int stall_left;
    int other;           // the earlier ticket that made this growth fail (-1: unknown)
    bool nonspec;        // at the commit head: everything before is committed
    bool have_pnext;
    unsigned int mine, pnext;
    float sumdx, sumdy, th, precdeg;
    double prec;
    unsigned int* ring;
    unsigned int* greg;  // the region list in global memory (nullptr while the region is small)
    unsigned int* areg;  // first half of the buffer (the first growth's list)
};

// ---------------------------------------------------------------------------------------------------------------
// region2rect() + get_theta(), refine()'s statistics, reduce_region_radius(): the reference's loops, one thread
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void lsd_rect_thread(const TEnv& E, const unsigned int* reg, int n, double reg_angle, double prec, double p, LsdRect& rec) {
    const int W = E.W;
    double x = 0, y = 0, sum = 0;
#pragma unroll 4
    for (int k = 0; k < n; k++) {
        const unsigned pp = reg[k];
        const int px = (int)(pp & 0xffffu), py = (int)(pp >> 16);
        PL_LSD_CHECK(px < W && py < E.H, 3);
        if (px >= W || py >= E.H) return;
        const double w = sqrt((double)E.g2[(size_t)py * W + px] / 4.0);  // modgrad
        x = __dadd_rn(x, __dmul_rn((double)px, w));
        y = __dadd_rn(y, __dmul_rn((double)py, w));
        sum = __dadd_rn(sum, w);
    }
    x = x / sum;
    y = y / sum;
    double Ixx = 0, Iyy = 0, Ixy = 0;
#pragma unroll 4
    for (int k = 0; k < n; k++) {
        const unsigned pp = reg[k];
        const int px = (int)(pp & 0xffffu), py = (int)(pp >> 16);
        const double w = sqrt((double)E.g2[(size_t)py * W + px] / 4.0);
        const double dx = __dsub_rn((double)px, x), dy = __dsub_rn((double)py, y);
        Ixx = __dadd_rn(Ixx, __dmul_rn(__dmul_rn(dy, dy), w));
        Iyy = __dadd_rn(Iyy, __dmul_rn(__dmul_rn(dx, dx), w));
        Ixy = __dsub_rn(Ixy, __dmul_rn(__dmul_rn(dx, dy), w));
    }
    const double dI = __dsub_rn(Ixx, Iyy);
    const double lambda = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ixx, Iyy), sqrt(__dadd_rn(__dmul_rn(dI, dI), __dmul_rn(__dmul_rn(4.0, Ixy), Ixy)))));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)__dsub_rn(lambda, Ixx), (float)Ixy)
                                           : (double)fast_atan2_deg((float)Ixy, (float)__dsub_rn(lambda, Iyy));
    theta *= kDegToRad;
    if (fabs(lsd_angle_diff_signed(theta, reg_angle)) > prec) theta += kPiD;
    double dx, dy;
    sincos(theta, &dy, &dx);
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
#pragma unroll 4
    for (int k = 0; k < n; k++) {
        const unsigned pp = reg[k];
        const double rdx = __dsub_rn((double)(int)(pp & 0xffffu), x), rdy = __dsub_rn((double)(int)(pp >> 16), y);
        const double l = __dadd_rn(__dmul_rn(rdx, dx), __dmul_rn(rdy, dy));
        const double w = __dadd_rn(__dmul_rn(-rdx, dy), __dmul_rn(rdy, dx));
        l_max = fmax(l_max, l); l_min = fmin(l_min, l);
        w_max = fmax(w_max, w); w_min = fmin(w_min, w);
    }
    rec.x1 = __dadd_rn(x, __dmul_rn(l_min, dx)); rec.y1 = __dadd_rn(y, __dmul_rn(l_min, dy));
    rec.x2 = __dadd_rn(x, __dmul_rn(l_max, dx)); rec.y2 = __dadd_rn(y, __dmul_rn(l_max, dy));
    rec.width = __dsub_rn(w_max, w_min);
    rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
    if (rec.width < 1.0) rec.width = 1.0;
}
__device__ __forceinline__ double lsd_density(int n, const LsdRect& rec) {
    return (double)n / (sqrt(lsd_dist_sq(rec.x1, rec.y1, rec.x2, rec.y2)) * rec.width);
}
// refine(): 2 * standard deviation of the level-line angles near the seed -> the tolerance of the re-growth
__device__ __noinline__ double lsd_refine_tau_thread(const TEnv& E, const unsigned int* reg, int n, const LsdRect& rec) {
    const int W = E.W;
    const unsigned p0 = reg[0];
    const int sx = (int)(p0 & 0xffffu), sy = (int)(p0 >> 16);
    const double xc = (double)sx, yc = (double)sy;
    const double ang_c = (double)E.ang[(size_t)sy * W + sx] * kDegToRad;
    double sum = 0, s_sum = 0;
    int cnt = 0;
#pragma unroll 2
    for (int k = 0; k < n; k++) {
        const unsigned pp = reg[k];
        const int px = (int)(pp & 0xffffu), py = (int)(pp >> 16);
        if (sqrt(lsd_dist_sq(xc, yc, (double)px, (double)py)) < rec.width) {
            const double ang_d = lsd_angle_diff_signed((double)E.ang[(size_t)py * W + px] * kDegToRad, ang_c);
            sum = __dadd_rn(sum, ang_d);
            s_sum = __dadd_rn(s_sum, __dmul_rn(ang_d, ang_d));
            cnt++;
        }
    }
    const double mean_angle = sum / (double)cnt;
    return 2.0 * sqrt(__dadd_rn(__dsub_rn(s_sum, __dmul_rn(__dmul_rn(2.0, mean_angle), sum)) / (double)cnt, __dmul_rn(mean_angle, mean_angle)));
}
// reduce_region_radius(): swap-with-last removal keeps the reference's point order.  Removed points give their stamp back, and
// the points that stay are stamped again with `mine_after` when it is over: a stamp of the region from before the reduction says
// nothing about membership any more, so whoever deferred to one must fail at commit.
__device__ __noinline__ bool lsd_reduce_thread(const TEnv& E, unsigned int* reg, int& n, unsigned int mine, unsigned int mine_after, double reg_angle,
                                               double prec, double p, LsdRect& rec, double density, double density_th) {
    const unsigned p0 = reg[0];
    const double xc = (double)(int)(p0 & 0xffffu), yc = (double)(int)(p0 >> 16);
    const double r1 = lsd_dist_sq(xc, yc, rec.x1, rec.y1), r2 = lsd_dist_sq(xc, yc, rec.x2, rec.y2);
    double radSq = r1 > r2 ? r1 : r2;
    int rounds = 0;
    while (density < density_th) {
        PL_LSD_CHECK(++rounds < 4096, 7);  // (the radius shrinks by a quarter every round)
        if (rounds >= 4096) break;
        radSq *= 0.75 * 0.75;
        int m = n;
        for (int k = 0; k < m; ++k) {
            const unsigned q = reg[k];
            const int qx = (int)(q & 0xffffu), qy = (int)(q >> 16);
            if (lsd_dist_sq(xc, yc, (double)qx, (double)qy) > radSq) {
                atomicCAS(&E.rec[(size_t)qy * E.W + qx].stamp, mine, 0u);
                reg[k] = reg[m - 1];
                reg[m - 1] = q;
                --m;
                --k;
            }
        }
        n = m;
        if (n < 2) break;
        lsd_rect_thread(E, reg, n, reg_angle, prec, p, rec);
        density = lsd_density(n, rec);
    }
    for (int k = 0; k < n; ++k) {
        const unsigned q = reg[k];
        atomicCAS(&E.rec[(size_t)(q >> 16) * E.W + (q & 0xffffu)].stamp, mine, mine_after);
    }
    return n >= 2;
}

// ---------------------------------------------------------------------------------------------------------------
// stamps
// ---------------------------------------------------------------------------------------------------------------
// What an earlier, uncommitted ticket's stamp `st` on a pixel means right now:
//   0 its growth is not final yet (or it failed: it will be grown again at its turn),  1 final, and the pixel is in its region,
//   2 final, and the region dropped the pixel (refine),  3 the ticket was committed meanwhile
__device__ __forceinline__ int lsd_owner_state(const TEnv& E, unsigned st) {
    const int u = (int)(st & kTicketMask) - 1;
    const unsigned w = *(volatile unsigned int*)&E.state[u & E.wm];
    const unsigned f = *(volatile unsigned int*)&E.fin[u & E.wm];
    if (*(volatile int*)&E.ctl->commit_head > u) return 3;  // (read last: the slot was still the ticket's when w and f were read)
    // a stamp of the committer's own growth (attempts 62, 63): the ticket at the commit head is being grown again right now — its
    // slot still shows the failed speculation, which must not be read as "finished, and the region dropped this pixel"
    if ((st >> 24) >= 62u) return 0;
    if ((w & 0xff) != kSlotDone || f == 0u) return 0;
    if (f != st && (E.debug & 4)) return 0;
    return f == st ? 1 : 2;
}
__device__ __forceinline__ bool lsd_add_dep(const TEnv& E, int ticket, unsigned st) {
    unsigned int* d = reinterpret_cast<unsigned int*>(&E.dep[ticket & E.wm]);
#pragma unroll
    for (int k = 0; k < kMaxDeps; k++) {
        const unsigned v = d[k];
        if (v == st) return true;
        if (v == 0u) {
            d[k] = st;
            if (E.prof) atomicAdd(&E.ctl->why[9], 1u);
            return true;
        }
    }
    return false;
}
// Writes this region's stamp on a pixel whose stamp was seen as `obs` (the caller decided the pixel may be taken).
// Speculative growers: compare-and-swap, false when somebody else stamped the pixel in between.  At the commit head: always.
// A later uncommitted ticket that loses the pixel is told so.
__device__ __forceinline__ bool lsd_stamp(const TEnv& E, const TState& T, int o, unsigned obs, int head) {
    unsigned old = obs;
    bool ok = true;
    if (T.nonspec) old = atomicExch(&E.rec[o].stamp, T.mine);
    else ok = atomicCAS(&E.rec[o].stamp, obs, T.mine) == obs;
    const unsigned tk = old & kTicketMask;
    if (ok && tk > (T.mine & kTicketMask) && (int)(tk - 1) >= head) atomicOr(&E.poison[(tk - 1) & E.wm], 1u << ((old >> 24) & 31u));
    return ok;
}

// a region reached min_reg_size: it gets a buffer in global memory
__device__ __noinline__ int lsd_thread_get_buffer(const TEnv& E, const unsigned int* ring, int n) {
    const int b = pool_pop_thread(E.ctl->free_mask);
    if (b < 0) return -1;
    unsigned int* dst = E.pool_reg + (size_t)b * 2 * kSpecCap;
#pragma unroll 1
    for (int k = 0; k < n; k++) dst[k] = ring[k];
    return b;
}

// Looks at the seed of T.ticket and starts the region (first growth, or refine's re-growth with T.prec / T.attempt set by the
// caller).  Returns false when the thread has to look again later (T.wait_u says for whom).
__device__ __forceinline__ bool lsd_thread_start(const TEnv& E, TState& T, int pix, bool first) {
    const volatile unsigned int* vused = E.used;
    if (first && ((vused[pix >> 5] >> (pix & 31)) & 1u)) {
        T.bad = 3;  // swallowed already: the committer voids it
        if (E.prof) atomicAdd(&E.ctl->why[0], 1u);
        return true;
    }
    const float4 r = lsd_ld_rec(E.rec + pix);
    const float2 c0 = E.cs0[pix];
    const unsigned obs = __float_as_uint(r.y);
    const unsigned tk = obs & kTicketMask, my_tk = T.mine & kTicketMask;
    const int head = *(volatile int*)&E.ctl->commit_head;
    if (tk != 0 && tk < my_tk && (int)(tk - 1) >= head) {  // an earlier uncommitted ticket holds the seed
        const int os = lsd_owner_state(E, obs);
        if (os == 3) return false;  // committed meanwhile: look at the map again
        if (os == 0 && (T.stall_left > 0 || ((obs >> 24) >= 62u && !(E.debug & 1024)))) {  // (the committer's own growth: it is committed in a moment)
            T.wait_u = (int)tk - 1;
            T.wait_hard = (obs >> 24) >= 62u && !(E.debug & 1024);
            return false;
        }
        if (os != 2) {  // most likely being swallowed by that region: the committer decides (void, or grown at its turn)
            T.bad = 3;
            if (E.prof) atomicAdd(&E.ctl->why[1], 1u);
            return true;
        }
    }
    if (!lsd_stamp(E, T, pix, obs, head)) {
        T.bad = 1;
        T.other = -1;
        if (E.prof) atomicAdd(&E.ctl->why[3], 1u);
        return true;
    }
    const unsigned pk = ((unsigned)(pix / E.W) << 16) | (unsigned)(pix % E.W);
    T.ring[0] = pk;
    if (T.greg) T.greg[0] = pk;
    T.n = 1;
    T.i = 0;
    T.kstart = 0;
    T.have_pnext = false;
    T.sumdx = c0.x;
    T.sumdy = c0.y;
    T.th = r.x;  // the region angle is the seed's until the first acceptance
    return true;
}

// The committed-map bits of (x - 1, x, x + 1) in row yy as bits 0..2 (1 = committed); pixels outside the image count as committed.
// (The bitmap has a spare word at its end.)
__device__ __forceinline__ unsigned lsd_used3(const volatile unsigned int* vused, int W, int H, int x, int yy) {
    if (yy < 0 || yy >= H) return 7u;
    const int o = yy * W + max(x - 1, 0);
    const unsigned s = (unsigned)o & 31u;
    unsigned bits = vused[o >> 5] >> s;
    if (s > 29u) bits |= vused[(o >> 5) + 1] << (32u - s);
    if (x == 0) bits = (bits << 1) | 1u;
    if (x + 1 >= W) bits |= 4u;
    return bits & 7u;
}

// One region point: its 8 neighbours in the reference's order (row-major, the centre skipped).
__device__ __forceinline__ void lsd_thread_step(const TEnv& E, TState& T) {
    const volatile unsigned int* vused = E.used;
    const int W = E.W, H = E.H;
    unsigned p;
    if (T.n - T.i <= kRing) p = T.ring[T.i & (kRing - 1)];
    else p = T.have_pnext ? T.pnext : T.greg[T.i];
    T.have_pnext = false;
    if (T.n - (T.i + 1) > kRing) {  // the point after this one has left the ring: its load goes out now
        T.pnext = T.greg[T.i + 1];
        T.have_pnext = true;
    }
    const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
    PL_LSD_CHECK(x < W && y < H, 1);
    if (x >= W || y >= H) {
        T.bad = 2;
        return;
    }
    const int o0 = y * W + x;
    const int head = *(volatile int*)&E.ctl->commit_head;
    const int ks = T.kstart;
    T.kstart = 0;
    float4 r[8];
    bool c[8];
    // committed bits of the 3 x 3 neighbourhood: bit 3 * (dy + 1) + (dx + 1)
    const unsigned ub = lsd_used3(vused, W, H, x, y - 1) | (lsd_used3(vused, W, H, x, y) << 3) | (lsd_used3(vused, W, H, x, y + 1) << 6);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int kk = k + (k >= 4 ? 1 : 0);
        const int dy = kk / 3 - 1, dx = kk - (kk / 3) * 3 - 1;
        const int o = o0 + dy * W + dx;
        c[k] = k >= ks && ((ub >> kk) & 1u) == 0;
        r[k] = make_float4(kNotDefDeg, 0.f, 0.f, 0.f);
        if (c[k]) r[k] = lsd_ld_rec(E.rec + o);
    }
    const unsigned my_tk = T.mine & kTicketMask;
    int casfail = 0;
    bool stop = false;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int kk = k + (k >= 4 ? 1 : 0);
        const int dy = kk / 3 - 1, dx = kk - (kk / 3) * 3 - 1;
        const float adeg = r[k].x;
        const unsigned st = __float_as_uint(r[k].y);
        if (c[k] && !stop && !T.bad && adeg != kNotDefDeg && st != T.mine && lsd_aligned_deg(T.th, adeg, T.precdeg, T.prec)) {
            const unsigned tk = st & kTicketMask;
            bool take = true;
            if (tk != 0 && tk < my_tk && (int)(tk - 1) >= head) {  // an earlier uncommitted ticket's stamp
                const int os = lsd_owner_state(E, st);
                if (os == 3) {  // committed meanwhile: look at this neighbour again (the map says what happened)
                    T.kstart = k;
                    stop = true;
                    take = false;
                } else if (os == 0 && (T.stall_left > 0 || ((st >> 24) >= 62u && !(E.debug & 1024)))) {  // (the committer's own growth: committed in a moment)
                    T.wait_u = (int)tk - 1;
                    T.wait_hard = (st >> 24) >= 62u && !(E.debug & 1024);
                    T.kstart = k;
                    stop = true;
                    take = false;
                    if (E.prof) atomicAdd(&E.ctl->why[6], 1u);
                } else if (os != 2) {  // the pixel is (assumed to stay) in that region: skipped as USED, checked at commit
                    take = false;
                    if (!lsd_add_dep(E, T.ticket, st)) {
                        T.bad = 3;
                        if (E.prof) atomicAdd(&E.ctl->why[2], 1u);
                    }
                }
            }
            if (take) {
                casfail |= !lsd_stamp(E, T, o0 + dy * W + dx, st, head);
                const unsigned pk = ((unsigned)(y + dy) << 16) | (unsigned)(x + dx);
                T.ring[T.n & (kRing - 1)] = pk;
                if (T.greg) T.greg[T.n] = pk;
                T.n++;
                if (T.n == E.min_reg && T.buf == -1) {
                    T.buf = lsd_thread_get_buffer(E, T.ring, T.n);
                    if (T.buf < 0) T.bad = 2;
                    else T.greg = T.areg = E.pool_reg + (size_t)T.buf * 2 * kSpecCap;
                } else if (T.n >= (T.buf != -2 ? kSpecCap : W * H)) {
                    PL_LSD_CHECK(T.buf != -2, 2);
                    T.bad = 2;
                }
                T.sumdx = __fadd_rn(T.sumdx, r[k].z);
                T.sumdy = __fadd_rn(T.sumdy, r[k].w);
                T.th = fast_atan2_deg(T.sumdy, T.sumdx);
            }
        }
    }
    if (casfail && !T.bad) {
        T.bad = 1;
        T.other = -1;
        if (E.prof) atomicAdd(&E.ctl->why[3], 1u);
    }
    if (!stop) T.i++;
    else T.have_pnext = false;  // (the prefetched point is the one after this one)
}

// The growth of T's region ended (or was given up).  Returns true when the ticket is finished: `status`, T.n, T.n0, n1, `rec` and
// the final list (T.greg, or the ring) describe it; false: refine grows the region again (the state is set up, keep stepping).
__device__ __noinline__ bool lsd_thread_growth_end(const TEnv& E, TState& T, int pix, int& status, LsdRect& rec, int& n1) {
    const double prec0 = kPiD * 22.5 / 180, p = 22.5 / 180, density_th = 0.7;
    if (!T.bad && !T.nonspec) {
        __threadfence_block();
        if ((*(volatile unsigned int*)&E.poison[T.ticket & E.wm] >> (T.attempt & 31)) & 1u) {  // an earlier ticket took one of its pixels
            T.bad = 1;
            T.other = -1;
        }
    }
    if (T.bad) {
        status = T.bad == 2 ? kStCapacity : kStDeferred;
        return true;
    }
    const double reg_angle = (double)T.th * kDegToRad;
    if ((T.attempt & 1) == 0) {
        if (T.n < E.min_reg) {
            status = kStNoRect;
            return true;
        }
        lsd_rect_thread(E, T.greg, T.n, reg_angle, prec0, p, rec);
        if (lsd_density(T.n, rec) >= density_th) {
            status = kStRect;
            return true;
        }
        // refine: grow the region again from its seed with the tolerance the angles near the seed suggest.  The pixels of the
        // first growth are "unmarked" by the new attempt number: a stamp of the old attempt counts as free.
        const double tau = lsd_refine_tau_thread(E, T.greg, T.n, rec);
        T.n0 = T.n;
        T.attempt++;
        T.mine = ((unsigned)T.attempt << 24) | ((unsigned)T.ticket + 1u);
        T.prec = tau;
        T.precdeg = (float)(tau * (180.0 / kPiD));
        T.greg = T.areg + (T.buf == -2 ? E.W * E.H : kSpecCap);  // the second half of the buffer
        T.wait_u = -1;
        T.wait_hard = false;
        T.stall_left = 0;  // (its seed carries its own old stamp: nothing to wait for)
        if (!lsd_thread_start(E, T, pix, false)) T.bad = 1;
        return false;
    }
    n1 = T.n;
    if (T.n < 2) {
        status = kStNoRect;
        return true;
    }
    lsd_rect_thread(E, T.greg, T.n, reg_angle, prec0, p, rec);
    const double density = lsd_density(T.n, rec);
    if (density >= density_th) {
        status = kStRect;
        return true;
    }
    const unsigned mine_after = T.mine | (32u << 24);
    status = lsd_reduce_thread(E, T.greg, T.n, T.mine, mine_after, reg_angle, prec0, p, rec, density, density_th) ? kStRect : kStNoRect;
    T.mine = mine_after;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// grower threads
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void lsd_grower_threads(const TEnv& E, int poll_ns) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    GrowCtl* ctl = E.ctl;
    volatile GrowCtl* vc = ctl;
    const int wm = E.wm;
    const int gw = (int)(threadIdx.x >> 5) - 2;  // grower warp number
    const double prec0 = kPiD * 22.5 / 180;
    TState T;
    T.ticket = -1;
    T.phase = 0;
    T.nonspec = false;
    T.n = T.i = T.n0 = 0;
    T.buf = -1;
    T.bad = 0;
    T.attempt = 0;
    T.kstart = 0;
    T.wait_u = -1;
    T.wait_hard = false;
    T.other = -1;
    T.stall_left = 0;
    T.restarts_left = 0;
    T.have_pnext = false;
    T.mine = T.pnext = 0;
    T.sumdx = T.sumdy = T.th = T.precdeg = 0.f;
    T.prec = prec0;
    T.ring = E.ring + (size_t)((int)threadIdx.x - 63) * kRing;  // (warps 0 and 1 are the committer and the issuer; ring 0 is the committer's)
    T.pix = 0;
    T.greg = T.areg = nullptr;
    long long busy = 0, t_take = 0, fit_cyc = 0;
    while (true) {
        int stop = 0;
        if (lane == 0) stop = vc->done | vc->abort_;
        if (__shfl_sync(FULL, stop, 0)) break;
        // ---- idle lanes take the next tickets (one atomic per warp) ----
        const unsigned idle = __ballot_sync(FULL, T.ticket < 0);
        if (idle) {
            int base = 0, cnt = 0;
            // Tickets are packed into the lower warps: a warp only takes new ones when the warps below it are three quarters busy,
            // so that a lightly loaded frame keeps few warps awake (a warp with one busy lane costs the SM as much as a full one).
            if (lane == 0 && (gw == 0 || vc->active >= 24 * gw)) {
                int gn = vc->grow_next;
                while (true) {
                    const int avail = vc->ticket_next - gn;
                    if (avail <= 0) break;
                    const int want = min(avail, __popc(idle));
                    const int old = atomicCAS(&ctl->grow_next, gn, gn + want);
                    if (old == gn) { base = gn; cnt = want; break; }
                    gn = old;
                }
                if (cnt) atomicAdd(&ctl->active, cnt);
            }
            base = __shfl_sync(FULL, base, 0);
            cnt = __shfl_sync(FULL, cnt, 0);
            if (cnt == 0 && idle == FULL) {
                __nanosleep(poll_ns);
                continue;
            }
            if (T.ticket < 0 && __popc(idle & lt) < cnt) {
                __threadfence_block();
                T.ticket = base + __popc(idle & lt);
                t_take = clock64();
                T.mine = (unsigned)T.ticket + 1u;
                T.n = T.i = T.n0 = 0;
                T.buf = -1;
                T.bad = 0;
                T.attempt = 0;
                T.greg = T.areg = nullptr;
                T.prec = prec0;
                T.precdeg = (float)(prec0 * (180.0 / kPiD));
                T.wait_u = -1;
                T.wait_hard = false;
                T.other = -1;
                T.stall_left = E.stall;
                T.restarts_left = E.restarts;
                // a seed right next to the previous ticket's seed most likely belongs to the same region: let that one go first
                T.phase = 1;
                const int sx = *(volatile int*)&E.hdr[T.ticket & wm].x;
                T.pix = sx & 0x7fffffff;
                T.kstart = sx < 0 ? 2 : 0;  // (in phase 1: iterations to wait)
            }
        }
        bool progress = false;
        if (T.ticket >= 0) {
            const int pix = T.pix;
            // ---- an earlier ticket took a pixel of this growth: it cannot be trusted ----
            if (!T.bad && T.phase == 2 && ((*(volatile unsigned int*)&E.poison[T.ticket & wm] >> (T.attempt & 31)) & 1u)) {
                T.bad = 1;
                T.other = -1;
            }
            // ---- waiting for an earlier ticket to become final ----
            if (T.wait_u >= 0) {
                const unsigned w = *(volatile unsigned int*)&E.state[T.wait_u & wm];
                if (vc->commit_head > T.wait_u || (!T.wait_hard && ((w & 0xff) == kSlotDone || --T.stall_left <= 0))) T.wait_u = -1;
            }
            if (T.wait_u < 0 && !T.bad) {
                progress = true;
                if (T.phase == 1) {
                    if (T.kstart > 0) T.kstart--;
                    else if (lsd_thread_start(E, T, pix, true)) T.phase = 2;
                } else if (T.i < T.n) {
                    lsd_thread_step(E, T);
                }
            }
            // ---- the growth is complete (or cannot go on) ----
            int status = 0, n1 = 0;
            bool finished = false;
            LsdRect rec;
            if (T.bad != 1 && (T.bad || (T.phase == 2 && T.i >= T.n && T.wait_u < 0))) {
                const long long f0 = clock64();
                finished = lsd_thread_growth_end(E, T, pix, status, rec, n1);
                fit_cyc += clock64() - f0;
                progress = true;
            }
            if (T.bad == 1) {
                if (T.restarts_left > 0) {
                    // start again from the seed as a new attempt (the stamps of the old one count as free), once the ticket that
                    // got in the way is final: its stamps are there to be seen now
                    if (E.prof) atomicAdd(&ctl->why[4], 1u);
                    T.restarts_left--;
                    T.attempt = (T.attempt & ~1) + 2;
                    T.mine = ((unsigned)T.attempt << 24) | ((unsigned)T.ticket + 1u);
                    E.dep[T.ticket & wm] = make_uint4(0u, 0u, 0u, 0u);
                    T.n = T.i = T.n0 = 0;
                    T.greg = T.buf >= 0 ? T.areg : nullptr;
                    T.bad = 0;
                    T.phase = 1;
                    T.prec = prec0;
                    T.precdeg = (float)(prec0 * (180.0 / kPiD));
                    T.wait_u = -1;
                    T.wait_hard = false;
                    T.stall_left = E.stall;
                    T.kstart = 4;  // (a few iterations for whoever got in the way to get on)
                    finished = false;
                } else {
                    status = kStDeferred;
                    finished = true;
                }
            }
            if (finished) {
                if (status < 0 && T.buf >= 0) {
                    pool_push(ctl->free_mask, T.buf);
                    T.buf = -1;
                }
                if (status == kStCapacity && E.prof) atomicAdd(&ctl->why[5], 1u);
                if (status == kStRect) E.pool_rect[T.buf] = rec;
                E.hdr[T.ticket & wm] = make_int4(pix, T.n, T.n0, n1);
                if (status >= 0 && T.buf < 0) {  // a small region: its points go with the ticket
                    E.pts4[T.ticket & wm] = make_uint4(T.ring[0], T.ring[1], T.ring[2], T.ring[3]);
                    if (T.n > 4) {
                        unsigned int* gp = E.trec[T.ticket & wm].pts;
#pragma unroll 1
                        for (int k = 0; k < T.n; k++) gp[k] = T.ring[k];
                    }
                }
                *(volatile unsigned int*)&E.fin[T.ticket & wm] = status >= 0 ? T.mine : 0u;
                __threadfence_block();
                *(volatile unsigned int*)&E.state[T.ticket & wm] = (unsigned)slot_pack(kSlotDone, status, T.buf);
                busy += clock64() - t_take;
                T.ticket = -1;
                atomicSub(&ctl->active, 1);
            }
        }
        if (!__any_sync(FULL, progress)) __nanosleep(poll_ns);
    }
    if (E.prof) {
        atomicAdd(&ctl->stat[4], (unsigned long long)busy);
        atomicAdd(&ctl->stat[5], (unsigned long long)fit_cyc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// issuer warp
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void lsd_issuer_warp(const TEnv& E, const GrowBufs& B, int f, int window, int poll_ns) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    volatile GrowCtl* vc = E.ctl;
    const volatile unsigned int* vused = E.used;
    const int wm = E.wm;
    const int ns = B.n_seeds[f];
    const unsigned int* sd = B.seeds + (size_t)f * B.plane;
    int t = 0;
    unsigned pix_next = lane < ns ? sd[lane] : 0u;
    for (int pos = 0; pos < ns; pos += 32) {
        const unsigned pix = pix_next;
        const int nxt = pos + 32 + lane;
        pix_next = nxt < ns ? sd[nxt] : 0u;
        const bool free_ = (pos + lane < ns) && ((vused[pix >> 5] >> (pix & 31)) & 1u) == 0;
        unsigned m = __ballot_sync(FULL, free_);
        // right neighbour of the previous seed of the chunk (same gradient bin, next pixel of the row): most likely the same region
        const unsigned prev = __shfl_up_sync(FULL, pix, 1);
        const bool adj = lane > 0 && ((m >> (lane - 1)) & 1u) && pix == prev + 1u;
        while (m) {
            int room = 0;
            while (true) {
                if (lane == 0) room = vc->abort_ ? -1 : window - (t - vc->commit_head);
                room = __shfl_sync(FULL, room, 0);
                if (room != 0) break;
                __nanosleep(poll_ns);
            }
            if (room < 0) return;
            const int take = min(room, __popc(m));
            const bool mine = ((m >> lane) & 1u) && __popc(m & lt) < take;
            if (mine) {
                const int tk = t + __popc(m & lt);
                E.poison[tk & wm] = 0u;
                E.fin[tk & wm] = 0u;
                E.hdr[tk & wm] = make_int4((int)(pix | (adj ? 0x80000000u : 0u)), 0, 0, 0);
                E.dep[tk & wm] = make_uint4(0u, 0u, 0u, 0u);
                E.state[tk & wm] = (unsigned)slot_pack(kSlotReady, 0, -1);
            }
            __syncwarp();
            __threadfence_block();
            t += take;
            if (lane == 0) vc->ticket_next = t;
            m &= ~__ballot_sync(FULL, mine);
        }
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) vc->all_issued = 1;
}

// ---------------------------------------------------------------------------------------------------------------
// test hook (PLSLAM_LSD_DEBUG bit 4): region_grow() of the ticket at the commit head, straight from the committed map with a
// private bitmap (no stamps), compared with the first growth of the speculation that is about to be committed
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void lsd_shadow_check(const TEnv& E, const GrowBufs& B, int f, int h, int pix, const unsigned int* spec, int n_spec,
                                              unsigned int* shadow, unsigned int* bits, unsigned fin_stamp, int n0) {
    const volatile unsigned int* vused = E.used;
    const int W = E.W, H = E.H;
    const double prec = kPiD * 22.5 / 180;
    const float precdeg = (float)(prec * (180.0 / kPiD));
    const float2 c0 = E.cs0[pix];
    float sumdx = c0.x, sumdy = c0.y, th = E.ang[pix];
    int n = 1;
    shadow[0] = ((unsigned)(pix / W) << 16) | (unsigned)(pix % W);
    bits[pix >> 5] |= 1u << (pix & 31);
    for (int i = 0; i < n; i++) {
        const unsigned p = shadow[i];
        const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
        for (int yy = max(y - 1, 0); yy <= min(y + 1, H - 1); yy++)
            for (int xx = max(x - 1, 0); xx <= min(x + 1, W - 1); xx++) {
                const int o = yy * W + xx;
                if ((vused[o >> 5] >> (o & 31)) & 1u) continue;
                if ((bits[o >> 5] >> (o & 31)) & 1u) continue;
                const float4 r = lsd_ld_rec(E.rec + o);
                if (r.x == kNotDefDeg || !lsd_aligned_deg(th, r.x, precdeg, prec)) continue;
                bits[o >> 5] |= 1u << (o & 31);
                shadow[n++] = ((unsigned)yy << 16) | (unsigned)xx;
                sumdx = __fadd_rn(sumdx, r.z);
                sumdy = __fadd_rn(sumdy, r.w);
                th = fast_atan2_deg(sumdy, sumdx);
            }
    }
    int diff = n != n_spec ? min(n, n_spec) : -1;
    for (int k = 0; k < min(n, n_spec); k++)
        if (shadow[k] != spec[k]) { diff = k; break; }
    for (int k = 0; k < n; k++) {
        const unsigned p = shadow[k];
        const int o = (int)(p >> 16) * W + (int)(p & 0xffffu);
        bits[o >> 5] &= ~(1u << (o & 31));
    }
    if (diff >= 0 && atomicCAS(B.dbg_out + (size_t)f * 16, 0, 1) == 0) {
        int* d = B.dbg_out + (size_t)f * 16;
        d[1] = h; d[2] = pix; d[3] = n_spec; d[4] = n; d[5] = diff; d[6] = (int)fin_stamp; d[7] = n0;
        d[8] = (int)E.poison[h & E.wm]; d[9] = 0;
        for (int k = 0; k < kMaxDeps; k++) d[10 + k] = (int)reinterpret_cast<const unsigned int*>(&E.dep[h & E.wm])[k];
        d[14] = diff < n_spec ? (int)spec[diff] : -1;
        d[15] = diff < n ? (int)shadow[diff] : -1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// committer warp
// ---------------------------------------------------------------------------------------------------------------
// were the stamps ticket `t` deferred to committed as they were seen?  (tickets below `h` are committed)
__device__ __forceinline__ bool lsd_deps_hold(const TEnv& E, int t, int h, const int* s_kind) {
    const uint4 dv = E.dep[t & E.wm];
    const unsigned int d[kMaxDeps] = {dv.x, dv.y, dv.z, dv.w};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < kMaxDeps; k++) {
        const unsigned st = d[k];
        if (st != 0u) {
            const int u = (int)(st & kTicketMask) - 1;
            if (u < h) ok &= *(volatile unsigned int*)&E.gfinal[u & (2 * E.wm + 1)] == st;
            else ok &= !(E.debug & 64) && *(volatile int*)&s_kind[u - h] == 1 && st == *(volatile unsigned int*)&E.fin[u & E.wm];  // a small region of the same batch
        }
    }
    return ok;
}

__device__ __noinline__ void lsd_committer_warp(const LineGeom& g, const TEnv& E, const GrowBufs& B, int f, int poll_ns, int* s_kind, LsdRect* s_rec) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    GrowCtl* ctl = E.ctl;
    volatile GrowCtl* vc = ctl;
    const volatile unsigned int* vused = E.used;
    const int wm = E.wm, W = E.W, fm = 2 * E.wm + 1;
    LsdQueueItem* q = B.queue + (size_t)f * g.seg_cap;
    unsigned int* my_big = B.big_reg + (size_t)blockIdx.x * 2 * B.plane;
    unsigned long long n_commit = 0, n_void = 0, n_regrown = 0, regrow_cyc = 0, wait_cyc = 0, batch_cyc = 0, solo_cyc = 0;
    int rects = 0;
    int h = 0;
    const long long t_begin = clock64();
    while (true) {
        int tn = 0, fin = 0;
        if (lane == 0) {
            fin = vc->all_issued;  // read before ticket_next: if it is set, ticket_next is final
            __threadfence_block();
            tn = vc->ticket_next;
            if (vc->abort_) fin = 2;
            else if (B.watchdog_cycles > 0 && clock64() - t_begin > B.watchdog_cycles) {
                vc->abort_ = 1;
                atomicOr(B.flags + f, 4);
                fin = 2;
            }
        }
        tn = __shfl_sync(FULL, tn, 0);
        fin = __shfl_sync(FULL, fin, 0);
        if (fin == 2) break;
        if (h == tn) {
            if (fin) break;  // every ticket is committed
            const long long w0 = clock64();
            __nanosleep(poll_ns);
            wait_cyc += (unsigned long long)(clock64() - w0);
            continue;
        }
        // ---- the run of finished tickets at the head, one per lane ----
        const int t = h + lane;
        unsigned w = 0;
        if (t < tn) w = *(volatile unsigned int*)&E.state[t & wm];
        const unsigned dm = __ballot_sync(FULL, t < tn && (w & 0xff) == kSlotDone);
        const int run = dm == FULL ? 32 : __ffs(~dm) - 1;
        if (run == 0) {
            const long long w0 = clock64();
            __nanosleep(100);
            wait_cyc += (unsigned long long)(clock64() - w0);
            continue;
        }
        __threadfence_block();
        const long long b0 = clock64();
        // kind: 0 void, 1 small region (no rectangle, no buffer), 3 region with a buffer, 4 failed speculation (void if an earlier region
        // swallows its seed, else grown again), 2 needs the ticket-alone path
        int kind = 2, n = 0, n0 = 0, nb = 0, buf = -1, status = 0, seed = 0;
        unsigned f1 = 0u;
        uint4 pv[kRing / 4];  // the points of a small region
        if (lane < run) {
            const int4 sv = E.hdr[t & wm];  // (stable once the ticket is done)
            seed = sv.x & 0x7fffffff;
            status = (int)((w >> 8) & 0xff) - 2;
            buf = (int)((w >> 16) & 0xffff) - 1;
            n = sv.y;
            n0 = sv.z;
            nb = sv.w;
            PL_LSD_CHECK(n >= 0 && (n <= kRing || buf >= 0 || status < 0), 6);
            f1 = *(volatile unsigned int*)&E.fin[t & wm];
            const bool poisoned = ((*(volatile unsigned int*)&E.poison[t & wm] >> ((f1 >> 24) & 31u)) & 1u) != 0u;
            if ((vused[seed >> 5] >> (seed & 31)) & 1u) kind = 0;
            else if (status < 0) kind = 4;
            else if (!poisoned && !(E.debug & 48)) {
                kind = buf < 0 ? 1 : 3;
                if (buf >= 0 && n + n0 + nb > 4096) kind = 2;  // (a long list: the ticket-alone path has nothing else to wait for it)
            }
        }
        s_kind[lane] = kind == 3 ? 1 : kind;  // (for a dependency inside the batch both commit the speculation they finished with)
        // the points of the small regions: four from the slot, the rest with vector loads from global memory, all issued together
        if (lane < run && kind == 1) {
            pv[0] = E.pts4[t & wm];
            const uint4* gp = reinterpret_cast<const uint4*>(E.trec[t & wm].pts);
#pragma unroll
            for (int j = 1; j < kRing / 4; j++)
                if (n > 4 * j) pv[j] = gp[j];
        }
        __syncwarp();
        // every pixel the growth ever accepted must still be uncommitted (the regions of one batch cannot share a pixel: a ticket that
        // lost one to an earlier ticket is poisoned), and the stamps it deferred to must have committed
        bool ok = true;
        if (lane < run && (kind == 1 || kind == 3)) ok = lsd_deps_hold(E, t, h, s_kind);
        if (lane < run && kind == 1 && ok) {
#pragma unroll
            for (int j = 0; j < kRing / 4; j++) {
                const unsigned q4[4] = {pv[j].x, pv[j].y, pv[j].z, pv[j].w};
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (4 * j + e < n) {
                        const unsigned o = (q4[e] >> 16) * (unsigned)W + (q4[e] & 0xffffu);
                        ok &= ((vused[o >> 5] >> (o & 31)) & 1u) == 0;
                    }
            }
        }
        {   // the lists of the regions with buffers, one region after the other with the whole warp
            unsigned bigm = __ballot_sync(FULL, lane < run && kind == 3 && ok);
            const unsigned stop_at = __ballot_sync(FULL, lane < run && (kind == 2 || ((kind == 1 || kind == 3) && !ok)));
            if (stop_at) bigm &= (1u << (__ffs(stop_at) - 1)) - 1u;  // (nothing behind the first failure commits now)
            while (bigm) {
                const int bl = __ffs(bigm) - 1;
                bigm &= bigm - 1;
                const int bbuf = __shfl_sync(FULL, buf, bl), bn = __shfl_sync(FULL, n, bl), bn0 = __shfl_sync(FULL, n0, bl), bnb = __shfl_sync(FULL, nb, bl);
                const unsigned int* rega = E.pool_reg + (size_t)bbuf * 2 * kSpecCap;
                const int la = bn0 > 0 ? bn0 : bn, lb = bn0 > 0 ? bnb : 0;
                bool conflict = false;
#pragma unroll 2
                for (int j = lane; j < la + lb; j += 32) {
                    const unsigned pp = j < la ? rega[j] : rega[kSpecCap + (j - la)];
                    const unsigned o = (pp >> 16) * (unsigned)W + (pp & 0xffffu);
                    PL_LSD_CHECK(o < (unsigned)(W * E.H), 4);
                    if (o < (unsigned)(W * E.H)) conflict |= ((vused[o >> 5] >> (o & 31)) & 1u) != 0;
                }
                if (__any_sync(FULL, conflict)) {
                    if (lane == bl) ok = false;
                    break;  // (what lies behind it does not commit in this batch)
                }
            }
        }
        if (lane < run && (kind == 1 || kind == 3) && !ok) kind = 2;
        // ---- commit: up to the first ticket that needs the ticket-alone path.  A failed speculation whose seed one of the regions
        //      committed here swallows is void and does not stop the batch ----
        int lo = 0, k = 0;
        while (true) {
            const unsigned stopm = __ballot_sync(FULL, lane >= lo && lane < run && (kind == 2 || kind == 4));
            k = stopm ? __ffs(stopm) - 1 : run;
            const bool mine = lane >= lo && lane < k;
            if (mine && kind == 1) {
#pragma unroll
                for (int j = 0; j < kRing / 4; j++) {
                    const unsigned q4[4] = {pv[j].x, pv[j].y, pv[j].z, pv[j].w};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (4 * j + e < n) {
                            const unsigned o = (q4[e] >> 16) * (unsigned)W + (q4[e] & 0xffffu);
                            const unsigned old = atomicOr(&E.used[o >> 5], 1u << (o & 31));
                            if (__builtin_expect((old >> (o & 31)) & 1u, 0)) atomicOr(B.flags + f, 8 | (8 << 4));  // two regions of a batch share a pixel: must not happen
                        }
                }
            }
            unsigned bigm = __ballot_sync(FULL, mine && kind == 3);
            while (bigm) {
                const int bl = __ffs(bigm) - 1;
                bigm &= bigm - 1;
                const int bbuf = __shfl_sync(FULL, buf, bl), bn = __shfl_sync(FULL, n, bl), bn0 = __shfl_sync(FULL, n0, bl), bst = __shfl_sync(FULL, status, bl);
                const unsigned int* rg = E.pool_reg + (size_t)bbuf * 2 * kSpecCap + (bn0 > 0 ? kSpecCap : 0);  // the final list
#pragma unroll 2
                for (int j = lane; j < bn; j += 32) {
                    const unsigned pp = rg[j];
                    const unsigned o = (pp >> 16) * (unsigned)W + (pp & 0xffffu);
                    PL_LSD_CHECK(o < (unsigned)(W * E.H), 5);
                    if (o < (unsigned)(W * E.H)) {
                        const unsigned old = atomicOr(&E.used[o >> 5], 1u << (o & 31));
                        if (__builtin_expect((old >> (o & 31)) & 1u, 0)) atomicOr(B.flags + f, 8 | (8 << 4));
                    }
                }
                if (bst == kStRect) {
                    if (rects < g.seg_cap) {
                        if (lane < (int)(sizeof(LsdRect) / 4))
                            reinterpret_cast<unsigned int*>(&q[rects].rec)[lane] = reinterpret_cast<const unsigned int*>(&E.pool_rect[bbuf])[lane];
                        rects++;
                    } else if (lane == 0) {
                        atomicOr(B.flags + f, 1);
                    }
                }
                if (lane == 0) pool_push(ctl->free_mask, bbuf);
            }
            if (mine) {
                E.gfinal[t & fm] = kind != 0 ? f1 : 0u;
                E.state[t & wm] = (unsigned)slot_pack(kSlotFree, 0, -1);
            }
            {
                const unsigned voids = __ballot_sync(FULL, mine && kind == 0);
                const unsigned regs = __ballot_sync(FULL, mine && kind != 0);
                if (B.dbg_log) {  // {seed, first growth size, final size, stamp | batch flag}
                    const int at = (int)n_commit + __popc(regs & ((1u << lane) - 1u));
                    if (mine && kind != 0 && at < kDbgLogCap)
                        B.dbg_log[(size_t)f * kDbgLogCap + at] = make_int4(seed, n0 > 0 ? n0 : n, n, (int)(f1 | 0x80000000u));
                }
                n_void += __popc(voids);
                n_commit += __popc(regs);
            }
            __syncwarp();
            if (k == run) break;
            // the ticket that stopped the batch: a failed speculation that the regions committed just now made void?
            bool now_void = false;
            if (lane == k && kind == 4) now_void = ((vused[seed >> 5] >> (seed & 31)) & 1u) != 0;
            if (!__any_sync(FULL, now_void)) break;
            if (lane == k) kind = 0;
            lo = k;
        }
        __threadfence_block();
        h += k;
        if (lane == 0) vc->commit_head = h;
        if (lane == 0 && E.prof) {
            ctl->why[10]++;
            batch_cyc += (unsigned long long)(clock64() - b0);
        }
        if (k == run) continue;
        const long long s0 = clock64();
        // ---- the ticket that ended the batch, alone ----
        {
            const int4 sv1 = E.hdr[h & wm];
            const unsigned w1 = *(volatile unsigned int*)&E.state[h & wm];
            const int pix = sv1.x & 0x7fffffff;
            int n1 = sv1.y, status = (int)((w1 >> 8) & 0xff) - 2;
            const int n0 = sv1.z, nb = sv1.w;
            const int buf = (int)((w1 >> 16) & 0xffff) - 1;
            const unsigned int* rega = buf >= 0 ? E.pool_reg + (size_t)buf * 2 * kSpecCap
                                                : (n1 > 4 ? E.trec[h & wm].pts : reinterpret_cast<const unsigned int*>(&E.pts4[h & wm]));
            const unsigned int* rg = n0 > 0 ? rega + kSpecCap : rega;  // the final list
            unsigned final_stamp = 0u;
            bool is_void = false;
            if ((vused[pix >> 5] >> (pix & 31)) & 1u) {
                is_void = true;  // swallowed by an earlier region
            } else {
                bool redo = status < 0;
                if (!redo) {  // is any pixel the growth ever accepted committed by now?  did the stamps it deferred to commit?
                    bool conflict = false;
                    const int la = n0 > 0 ? n0 : n1, lb = n0 > 0 ? nb : 0;
#pragma unroll 1
                    for (int j = lane; j < la + lb; j += 32) {
                        const unsigned pp = j < la ? rega[j] : rega[kSpecCap + (j - la)];
                        const unsigned o = (pp >> 16) * (unsigned)W + (pp & 0xffffu);
                        PL_LSD_CHECK(o < (unsigned)(W * E.H), 4);
                        if (o >= (unsigned)(W * E.H)) continue;
                        conflict |= ((vused[o >> 5] >> (o & 31)) & 1u) != 0;
                    }
                    const bool deps_ok = lsd_deps_hold(E, h, h, s_kind);
                    redo = __any_sync(FULL, conflict) || !deps_ok;
                    if (redo && lane == 0 && E.prof) atomicAdd(&ctl->why[deps_ok ? 7 : 8], 1u);
                }
                const bool regrown = redo;
                if (redo) {  // everything before this ticket is committed: this growth is the sequential one (one thread)
                    const long long g0 = clock64();
                    int nn = 0, use_b = 0;
                    if (lane == 0) {
                        TState T;
                        T.ticket = h;
                        T.phase = 2;
                        T.nonspec = true;
                        T.n = T.i = T.n0 = 0;
                        T.buf = -2;
                        T.bad = 0;
                        T.attempt = 62;
                        T.kstart = 0;
                        T.wait_u = -1;
                        T.wait_hard = false;
                        T.other = -1;
                        T.stall_left = 0;
                        T.restarts_left = 0;
                        T.have_pnext = false;
                        T.mine = (62u << 24) | ((unsigned)h + 1u);
                        T.pnext = 0;
                        T.prec = kPiD * 22.5 / 180;
                        T.precdeg = (float)(T.prec * (180.0 / kPiD));
                        T.ring = E.ring;  // (ring 0 is the committer's)
                        T.pix = pix;
                        T.greg = T.areg = my_big;
                        T.sumdx = T.sumdy = T.th = 0.f;
                        LsdRect rec;
                        int nb2 = 0;
                        lsd_thread_start(E, T, pix, false);  // (at the head nothing earlier is uncommitted: it never has to wait)
                        while (true) {
                            int guard = 0;
                            while (T.i < T.n && !T.bad) {
                                lsd_thread_step(E, T);
                                if ((++guard & 0xfff) == 0 && (vc->abort_ || (B.watchdog_cycles > 0 && clock64() - t_begin > B.watchdog_cycles))) T.bad = 2;
                            }
                            if (lsd_thread_growth_end(E, T, pix, status, rec, nb2)) break;
                        }
                        if (status == kStRect) *s_rec = rec;
                        nn = T.n;
                        use_b = T.n0 > 0;
                        if ((E.debug & 512) && status >= 0)
                            lsd_shadow_check(E, B, f, h, pix, my_big, T.n0 > 0 ? T.n0 : T.n, B.dbg_scratch + (size_t)blockIdx.x * B.plane,
                                             B.dbg_bits + (size_t)blockIdx.x * ((E.W * E.H + 31) / 32), T.mine, T.n0);
                    }
                    status = __shfl_sync(FULL, status, 0);
                    n1 = __shfl_sync(FULL, nn, 0);
                    use_b = __shfl_sync(FULL, use_b, 0);
                    rg = my_big + (use_b ? B.plane : 0);
                    if (status < 0 && lane == 0) atomicOr(B.flags + f, 2);
                    n_regrown++;
                    regrow_cyc += (unsigned long long)(clock64() - g0);
                    __syncwarp();
                } else {
                    final_stamp = *(volatile unsigned int*)&E.fin[h & wm];
                    if (((E.debug & 16) || ((E.debug & 128) && buf >= 0)) && lane == 0)
                        lsd_shadow_check(E, B, f, h, pix, rega, n0 > 0 ? n0 : n1, my_big, B.dbg_bits + (size_t)blockIdx.x * ((E.W * E.H + 31) / 32), final_stamp, n0);
                    __syncwarp();
                }
                if (status >= 0) {
#pragma unroll 1
                    for (int j = lane; j < n1; j += 32) {
                        const unsigned pp = rg[j];
                        const unsigned o = (pp >> 16) * (unsigned)W + (pp & 0xffffu);
                        PL_LSD_CHECK(o < (unsigned)(W * E.H), 5);
                        if (o >= (unsigned)(W * E.H)) continue;
                        atomicOr(&E.used[o >> 5], 1u << (o & 31));
                    }
                    if (status == kStRect) {
                        if (rects < g.seg_cap) {
                            const unsigned int* src = reinterpret_cast<const unsigned int*>(regrown ? s_rec : &E.pool_rect[buf]);
                            if (lane < (int)(sizeof(LsdRect) / 4)) reinterpret_cast<unsigned int*>(&q[rects].rec)[lane] = src[lane];
                            rects++;
                        } else if (lane == 0) {
                            atomicOr(B.flags + f, 1);
                        }
                    }
                }
            }
            if (!is_void && B.dbg_log && lane == 0 && n_commit < kDbgLogCap)
                B.dbg_log[(size_t)f * kDbgLogCap + n_commit] =
                    make_int4(pix, final_stamp ? (n0 > 0 ? n0 : n1) : -1, status >= 0 ? n1 : -1,
                              (int)(final_stamp | (*(volatile unsigned int*)&E.poison[h & wm] ? 0x40000000u : 0u)));
            if (is_void) n_void++;
            else n_commit++;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) {
                E.gfinal[h & fm] = final_stamp;
                if (buf >= 0) pool_push(ctl->free_mask, buf);
                E.state[h & wm] = (unsigned)slot_pack(kSlotFree, 0, -1);
                __threadfence_block();
                vc->commit_head = h + 1;
            }
            h++;
            __syncwarp();
            if (lane == 0 && E.prof) {
                ctl->why[11]++;
                solo_cyc += (unsigned long long)(clock64() - s0);
            }
        }
    }
    if (lane == 0) {
        B.n_rects[f] = rects;
        if (B.dbg_log_n) B.dbg_log_n[f] = (int)n_commit;
        ctl->stat[0] = n_commit;
        ctl->stat[1] = n_void;
        ctl->stat[2] = n_regrown;
        ctl->stat[3] = wait_cyc;
        ctl->stat[6] = (unsigned long long)(clock64() - t_begin);
        ctl->stat[7] = regrow_cyc;
        ctl->why[12] = (unsigned)(batch_cyc >> 10);
        ctl->why[13] = (unsigned)(solo_cyc >> 10);
        __threadfence_block();
        vc->done = 1;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel: warp 0 commits, warp 1 issues, the threads of all the other warps grow regions
// ---------------------------------------------------------------------------------------------------------------
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_lsd_grow(LineGeom g, GrowLayout L, int nf, GrowBufs B) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ GrowCtl s_ctl;
    __shared__ LsdRect s_rec;
    __shared__ int s_kind[32];
    __shared__ int s_frame;
    const int tid = threadIdx.x, warp = tid >> 5;
    TEnv E;
    E.ctl = &s_ctl;
    E.hdr = reinterpret_cast<int4*>(s_raw + L.off_hdr());
    E.dep = reinterpret_cast<uint4*>(s_raw + L.off_dep());
    E.pts4 = reinterpret_cast<uint4*>(s_raw + L.off_pts4());
    E.state = reinterpret_cast<unsigned int*>(s_raw + L.off_state());
    E.fin = reinterpret_cast<unsigned int*>(s_raw + L.off_fin());
    E.poison = reinterpret_cast<unsigned int*>(s_raw + L.off_poison());
    E.ring = reinterpret_cast<unsigned int*>(s_raw + L.off_ring());
    E.trec = B.trec + (size_t)blockIdx.x * kMaxSlots;
    E.gfinal = B.gfinal + (size_t)blockIdx.x * 2 * kMaxSlots;
    E.used = reinterpret_cast<unsigned int*>(s_raw + L.off_used());
    E.wm = L.W - 1;
    E.W = g.W;
    E.H = g.H;
    E.min_reg = g.min_reg_size;
    E.stall = L.stall;
    E.restarts = L.restarts;
    E.debug = L.debug;
    E.pool_reg = B.pool_reg + (size_t)blockIdx.x * kPoolBufs * 2 * kSpecCap;
    E.pool_rect = B.pool_rect + (size_t)blockIdx.x * kPoolBufs;
    E.prof = B.phase_cycles != nullptr;
    while (true) {
        if (tid == 0) s_frame = atomicAdd(B.frame_counter, 1);
        __syncthreads();
        const int f = s_frame;
        if (f >= nf) break;
        E.rec = B.rec + (size_t)f * B.plane;
        E.cs0 = B.cs0 + (size_t)f * B.plane;
        E.ang = B.angdeg + (size_t)f * B.plane;
        E.g2 = B.g2 + (size_t)f * B.plane;
        E.flag = B.flags + f;
        for (int k = tid; k < L.W; k += kThreads) {
            E.state[k] = (unsigned)slot_pack(kSlotFree, 0, -1);
            E.poison[k] = 0u;
            E.fin[k] = 0u;
            E.gfinal[k] = 0u;
            E.gfinal[k + L.W] = 0u;
        }
        for (int k = tid; k < L.bits_words + 1; k += kThreads) E.used[k] = 0u;
        if (tid == 0) {
            s_ctl.ticket_next = s_ctl.grow_next = s_ctl.commit_head = 0;
            s_ctl.all_issued = s_ctl.done = s_ctl.abort_ = 0;
            s_ctl.active = 0;
            for (int k = 0; k < kPoolBufs / 64; k++) s_ctl.free_mask[k] = ~0ull;
            for (int k = 0; k < 8; k++) s_ctl.stat[k] = 0;
            for (int k = 0; k < 16; k++) s_ctl.why[k] = 0;
            s_ctl.t_start = clock64();
        }
        __syncthreads();
        if (warp == 0) lsd_committer_warp(g, E, B, f, L.poll_ns, s_kind, &s_rec);
        else if (warp == 1) lsd_issuer_warp(E, B, f, L.window, L.poll_ns);
        else lsd_grower_threads(E, L.poll_ns);
        __syncthreads();
        if (tid == 0 && B.phase_cycles) {
            long long* pc = B.phase_cycles + (size_t)f * 24;
            pc[0] = (long long)s_ctl.stat[4];   // cycles grower threads held a ticket (summed over threads)
            pc[1] = clock64() - s_ctl.t_start;  // cycles the frame took
            pc[2] = (long long)s_ctl.stat[6];   // cycles of the committer warp (including waiting for the head ticket)
            pc[3] = s_ctl.ticket_next;
            pc[4] = (long long)s_ctl.stat[2];   // tickets grown again at commit time
            pc[5] = (long long)s_ctl.stat[0];   // committed regions
            pc[6] = (long long)s_ctl.stat[5];   // cycles grower threads spent fitting rectangles (region2rect, refine)
            pc[7] = (long long)s_ctl.stat[1];   // void tickets
            pc[8] = (long long)s_ctl.stat[7];   // cycles of the re-growths at commit time
            for (int k = 0; k < 10; k++) pc[9 + k] = s_ctl.why[k];
            pc[19] = s_ctl.why[10];             // batch iterations of the committer
            pc[20] = s_ctl.why[11];             // tickets committed alone
            pc[22] = s_ctl.why[12];             // kcycles of the batch iterations
            pc[23] = s_ctl.why[13];             // kcycles of the tickets committed alone
            pc[21] = (long long)s_ctl.stat[3];  // cycles the committer slept waiting for the head ticket
        }
        __syncthreads();
    }
}

}  // namespace pl
