// lsd_grow2.cuh — k_lsd_grow2: the ordered region growing of LSD (flsd() main loop, lsd.cpp; called from the reference at
// src/LineExtractor.cpp:21) as speculative transactions with in-order commit, one frame per CTA, warps specialised by role.
// (Included by line_kernels.cu, which defines the region-level functions: lsd_region_grow, lsd_region2rect, lsd_refine ...)
//
//   warp 0 — the SEQUENCER of the frame: everything that is serial by definition lives here, without locks.  It walks the seed
//            list (32 seeds per step, four chunks of the list in flight), gives every seed that is unused in the committed USED
//            map the next ticket, and commits finished tickets strictly in ticket order: seed already committed -> void; a
//            logged pixel already committed -> the growth depended on a region committed after the grower read the map, so the
//            sequencer grows the seed itself, now, when everything before it is committed (this growth IS the sequential one);
//            otherwise the speculative growth is what the sequential algorithm would have done: the region is OR-ed into the
//            committed map and its rectangle queued for NFA validation.
//   warps 1.. — GROWERS: take the next ticket (one atomicAdd), grow it against the committed map with private marks, fit and
//            refine the rectangle, park the result in the ticket's slot (tiny regions never leave shared memory).
//
// The rules that make the result the reference's are those of DESIGN.md 4.1: committed pixels are never released, tickets follow
// the seed order, a grower only ever sees pixels of earlier tickets in the map, and every pixel a grower ever accepted is in its
// log, which is checked against the map at commit time.  The claim stamps are a hint that saves wasted growth.
//
// Several CTAs share an SM (the committed bitmap of a 640x480 frame is 24.5 KB): while one frame's growers wait for the in-order
// head, another frame's growers run.  The role split keeps each warp's hot code small (the old single-loop kernel was bound by
// instruction fetch: 170 KB of SASS against a 32 KB L1.5 instruction cache).
#pragma once

namespace pl {

constexpr int kSlots2 = 128;     // ticket slots per frame: the window of uncommitted tickets is at most this
constexpr int kSpecCap2 = 8192;  // region / log capacity of a speculative grower (larger regions are grown by the sequencer)
constexpr int kMaxPool2 = 64;    // region buffers per CTA (bit mask)

struct Grow2Smem {
    int tiles, pool_tiles;
    int window;      // tickets that may be uncommitted at once (<= kSlots2)
    int lookahead;   // tickets issued ahead of the growers' demand (issuing late = fresher map = fewer void growths)
    int bits_words;  // words of a W*H bitmap
    int tail_nfa, poll_ns;
    int pool_n;      // region buffers of this CTA
    int split;       // 1: warp 0 commits, warp 1 issues tickets, growers from warp 2 on; 0: warp 0 does both, growers from warp 1 on
    // per warp: sval | ring | scratch (3 x 32 doubles) | frame view | result | tile pool | rev | dir | ntiles
    // (the sequencer re-grows with the full bitmap: no tile pool)
    __host__ __device__ size_t off_ring() const { return kSvalEntries * sizeof(float2); }
    __host__ __device__ size_t off_scratch() const { return off_ring() + kRegRing * sizeof(unsigned int); }
    __host__ __device__ size_t off_view() const { return off_scratch() + 96 * sizeof(double); }
    __host__ __device__ size_t off_res() const { return off_view() + ((sizeof(LsdFrame) + 15) & ~(size_t)15); }
    __host__ __device__ size_t off_pool() const { return off_res() + ((sizeof(GrowResult) + 15) & ~(size_t)15); }
    __host__ __device__ size_t off_rev() const { return off_pool() + (size_t)pool_tiles * 32 * sizeof(unsigned int); }
    __host__ __device__ size_t off_dir() const { return off_rev() + (((size_t)pool_tiles * sizeof(unsigned short) + 15) & ~(size_t)15); }
    __host__ __device__ size_t off_ntiles() const { return off_dir() + (((size_t)tiles + 15) & ~(size_t)15); }
    __host__ __device__ size_t per_grower() const { return off_ntiles() + 16; }
    __host__ __device__ size_t seq_area() const { return off_pool(); }
    // per CTA: ticket slots | committed bitmap | tiny regions | sequencer area | grower areas
    __host__ __device__ size_t off_used() const { return kSlots2 * sizeof(int4); }
    __host__ __device__ size_t off_tiny() const { return off_used() + (((size_t)bits_words * sizeof(unsigned int) + 15) & ~(size_t)15); }
    __host__ __device__ size_t off_seq() const { return off_tiny() + (size_t)kSlots2 * kTiny * sizeof(unsigned int); }
    __host__ __device__ size_t off_growers() const { return off_seq() + seq_area(); }
    __host__ __device__ size_t total(int warps) const { return off_growers() + (size_t)(warps - 1) * per_grower(); }
};
// ticket slot (shared memory, int4): x = seed pixel, y = state | (status + 2) << 8 | (buffer + 1) << 16, z = region size, w = log size
struct Grow2Ctl {
    int ticket_next;  // tickets issued           (sequencer writes)
    int grow_next;    // tickets claimed          (growers: atomicAdd)
    int commit_head;  // tickets committed        (sequencer writes)
    int frame_done;
    int all_issued;   // the seed list is exhausted (the issuer, when it is a warp of its own)
    int frame, ns;
    unsigned long long free_mask;  // free region buffers
};

// What the roles share, in static shared memory at namespace scope: the role functions are compiled out of line (so that each
// gets its own register allocation: inlined into one body, the sequencer's loop state lived in local memory) and read what they
// need from here with plain shared-memory loads.
struct Grow2Shared {
    Grow2Ctl ctl;
    LineGeom g;
    Grow2Smem gs;
    GrowBufs B;
    int nf;
    unsigned long long stat[8];  // profiling: growers' cycles growing, waiting for a ticket, parking, given up | sequencer's
    struct Seq {                 // the sequencer's state while the kernel body runs a commit-time re-growth for it
        int t_next, head, base, n_rect, n_commit, n_void, n_regrow, n_defer, all_issued;
        unsigned done_mask;
        long long c_commit, c_regrow, c_issue, c_idle, t_start, g0;
    } seq;
};
__shared__ Grow2Shared g2s;

// A full chunk of kNfaChunk committed rectangles of frame f goes to the queue of the tail helpers (bit 31: the chunk is full, its
// size does not depend on the frame's final count).  Whole warp: the rectangle words were written by several lanes.
__device__ __forceinline__ void g2_publish_chunk(const GrowBufs& B, int f, int chunk) {
    if (chunk >= kNfaChunksPerFrame) return;
    __threadfence();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const int b = atomicAdd(B.nfa_ctl + 0, 1);
        *(volatile unsigned int*)(B.nfa_items + b) = 0x80000000u | ((unsigned)f * kNfaChunksPerFrame + (unsigned)chunk);
    }
    __syncwarp();
}

// =============================== sequencer (warp 0) ===============================
// Returns -1 when the frame is finished, or the seed pixel of the head ticket when that ticket has to be grown now (deferred,
// over a capacity, or in conflict with a region committed after its grower read the map): the kernel body calls lsd_grow_seed
// with view[0] and comes back with resume = 1.  (The call is made there and not here so that this function contains no call:
// with one, its loop state lived in local memory.)
template <bool kProf, bool kIssue>
__device__ __noinline__ int g2_sequencer(int f, int resume) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    const LineGeom& g = g2s.g;
    const Grow2Smem& gs = g2s.gs;
    const GrowBufs& B = g2s.B;
    const size_t plane = B.plane;
    volatile int4* s_slot = reinterpret_cast<volatile int4*>(s_raw);
    unsigned int* s_used = reinterpret_cast<unsigned int*>(s_raw + gs.off_used());
    const volatile unsigned int* vused = s_used;
    unsigned int* s_tiny = reinterpret_cast<unsigned int*>(s_raw + gs.off_tiny());
    volatile Grow2Ctl* ctl = &g2s.ctl;
    Grow2Ctl& s_ctl = g2s.ctl;
    LsdFrame* s_view = reinterpret_cast<LsdFrame*>(s_raw + gs.off_seq() + gs.off_view());   // [0]: the sequencer's own
    GrowResult* s_res = reinterpret_cast<GrowResult*>(s_raw + gs.off_seq() + gs.off_res());
    unsigned int* my_pool_reg = B.pool_reg + (size_t)blockIdx.x * gs.pool_n * kSpecCap2;
    unsigned int* my_pool_touched = B.pool_touched + (size_t)blockIdx.x * gs.pool_n * kSpecCap2;
    LsdRect* my_pool_rect = B.pool_rect + (size_t)blockIdx.x * gs.pool_n;
    unsigned int* my_small = B.small_buf + (size_t)blockIdx.x * kSlots2 * 2 * kSmall;
    LsdRect* my_small_rect = B.small_rect + (size_t)blockIdx.x * kSlots2;
    Grow2Shared::Seq& Q = g2s.seq;
    const long long t_start = resume ? Q.t_start : clock64();
    const int ns = s_ctl.ns;
    const unsigned int* sd = B.seeds + (size_t)f * plane;
    LsdQueueItem* q = B.queue + (size_t)f * g.seg_cap;
    int t_next = 0, head = 0, base = 0, n_rect = 0;
    int n_commit = 0, n_void = 0, n_regrow = 0, n_defer = 0;
    long long c_commit = 0, c_regrow = 0, c_issue = 0, c_idle = 0;
    constexpr bool prof = kProf;  // the clock64 accounting of pl_line_grow_phases is compiled into a copy of its own
    unsigned done_mask = 0;
    bool all_issued = kIssue ? ns == 0 : false;
    // rectangles go to the validation queue of the tail helpers in chunks AS THEY ARE COMMITTED (not when the frame ends): CTAs that
    // have finished their frame then find work from every frame that is still growing
    const bool early_nfa = gs.tail_nfa != 0 && g2s.nf > 1;
    if (resume) {
        t_next = Q.t_next; head = Q.head; base = Q.base; n_rect = Q.n_rect;
        n_commit = Q.n_commit; n_void = Q.n_void; n_regrow = Q.n_regrow; n_defer = Q.n_defer;
        c_commit = Q.c_commit; c_regrow = Q.c_regrow; c_issue = Q.c_issue; c_idle = Q.c_idle;
        done_mask = Q.done_mask;
        all_issued = Q.all_issued != 0;
    }
    unsigned pa = 0, pb = 0, pc = 0, pd = 0;
    if (kIssue) {
        pa = base + lane < ns ? sd[base + lane] : 0;
        pb = base + 32 + lane < ns ? sd[base + 32 + lane] : 0;
        pc = base + 64 + lane < ns ? sd[base + 64 + lane] : 0;
        pd = base + 96 + lane < ns ? sd[base + 96 + lane] : 0;
    }
    if (resume) {
        // the head ticket has just been grown by the kernel body (everything before it is committed: that growth is the sequential
        // one): write it to the map, queue its rectangle
        const int status = s_res[0].status, n = s_res[0].n;
        const unsigned int* rg = B.big_reg + (size_t)f * plane;
        if (status < 0 && lane == 0) atomicOr(B.flags + f, 2);
        if (status >= 0) {
            #pragma unroll 1
            for (int i = lane; i < n; i += 32) {
                const unsigned pp = rg[i];
                const unsigned o = (pp >> 16) * (unsigned)g.W + (pp & 0xffffu);
                atomicOr(&s_used[o >> 5], 1u << (o & 31));
            }
            if (status == kStRect) {
                if (n_rect < g.seg_cap) {
                    if (lane < (int)(sizeof(LsdRect) / 4))
                        reinterpret_cast<unsigned int*>(&q[n_rect].rec)[lane] = reinterpret_cast<const unsigned int*>(&s_res[0].rec)[lane];
                    n_rect++;
                    if (early_nfa && (n_rect & (kNfaChunk - 1)) == 0) g2_publish_chunk(B, f, n_rect / kNfaChunk - 1);
                } else if (lane == 0) {
                    atomicOr(B.flags + f, 1);
                }
            }
        }
        n_commit++;
        n_regrow++;
        __syncwarp();
        __threadfence_block();
        head++;
        if (lane == 0) ctl->commit_head = head;
        if (prof) {
            const long long dt = clock64() - Q.g0;
            c_regrow += dt;
            c_commit += dt;
        }
    }
    if (!resume && lane == 0) {  // the view of the commit-time re-growth: no size limits, full private bitmap in global memory
        unsigned char* s_mine = s_raw + gs.off_seq();
        LsdFrame F;
        F.sval = reinterpret_cast<float2*>(s_mine);
        F.ring = reinterpret_cast<unsigned int*>(s_mine + gs.off_ring());
        F.scratch = reinterpret_cast<double*>(s_mine + gs.off_scratch());
        F.pool = nullptr; F.rev = nullptr; F.dir = nullptr; F.ntiles = nullptr;
        F.tw = (g.W + 31) >> 5;
        F.pool_tiles = 0;
        F.sparse = false;
        F.bits = B.big_bits + (size_t)blockIdx.x * gs.bits_words;
        F.ticket = 0;
        F.W = g.W;
        F.H = g.H;
        F.reg = B.big_reg + (size_t)f * plane;
        F.touched = B.big_touched + (size_t)blockIdx.x * 2 * plane;
        F.touched_buf = F.touched;
        F.reg_cap = (int)plane;
        F.touched_cap = (int)(2 * plane);
        F.used_bits = s_used;
        F.commit_head = &g2s.ctl.commit_head;
        F.ang = B.angdeg + (size_t)f * plane;
        F.g2 = B.g2 + (size_t)f * plane;
        F.rec = B.rec + (size_t)f * plane;
        F.cs0 = B.cs0 + (size_t)f * plane;
        s_view[0] = F;
    }
    __syncwarp();
    while (true) {
        bool progress = false;
        if (!kIssue) t_next = ctl->ticket_next;  // (a slot read before its ticket's state is visible says Free: see the end of a run)
        // ---------------- commit the run of finished tickets at the head ----------------
        // The states of up to 32 slots are read at once and the fences are paid once per run: when the growers are ahead of the
        // sequencer (the case that matters) a ticket costs a few shared-memory round trips.
        if (head < t_next) {
            int st = 0;
            if (lane < t_next - head) st = s_slot[(head + lane) & (kSlots2 - 1)].y;
            const unsigned dm = __ballot_sync(FULL, (st & 0xff) == kSlotDone);
            const int run = dm == FULL ? 32 : __ffs(~dm) - 1;
            if (run > 0) {
                const long long c0 = prof ? clock64() : 0;
                __threadfence_block();
                int my_pix = 0, my_n = 0, my_nt = 0;
                if (lane < run) {
                    volatile int4* sl = &s_slot[(head + lane) & (kSlots2 - 1)];
                    my_pix = sl->x;
                    my_n = sl->z;
                    my_nt = sl->w;
                }
                #pragma unroll 1
                for (int r = 0; r < run; r++) {
                    const int slot = (head + r) & (kSlots2 - 1);
                    const int w = __shfl_sync(FULL, st, r), pix = __shfl_sync(FULL, my_pix, r);
                    int n = __shfl_sync(FULL, my_n, r), nt = __shfl_sync(FULL, my_nt, r), status = ((w >> 8) & 0xff) - 2;
                    const int buf = ((w >> 16) & 0xff) - 1;
                    if (status == kStDeferred) n_defer++;
                    if ((vused[pix >> 5] >> (pix & 31)) & 1u) {
                        n_void++;  // swallowed by an earlier region
                        if (buf >= 0 && lane == 0) atomicOr(&s_ctl.free_mask, 1ull << buf);
                        continue;
                    }
                    const unsigned int* rg = buf >= 0 ? my_pool_reg + (size_t)buf * kSpecCap2 : my_small + (size_t)slot * 2 * kSmall;
                    bool redo = status < 0;
                    unsigned rg0 = 0, tk0 = 0, rc0 = 0;
                    // Most tickets are regions of a few pixels without a rectangle: their points never left shared memory
                    // (the grower parked them in the ticket slot) and their accept log is their point list.
                    const bool tiny = !redo && status == kStNoRect && buf < 0 && n <= kTiny && nt == 0;
                    if (tiny) {
                        bool conflict = false;
                        if (lane < n) {
                            rg0 = s_tiny[slot * kTiny + lane];
                            const unsigned o = (rg0 >> 16) * (unsigned)g.W + (rg0 & 0xffffu);
                            conflict = ((vused[o >> 5] >> (o & 31)) & 1u) != 0;
                        }
                        redo = __any_sync(FULL, conflict);
                    } else if (!redo) {
                        // points, log and rectangle live in global memory: the three reads are issued together
                        const unsigned int* tk = buf >= 0 ? my_pool_touched + (size_t)buf * kSpecCap2 : rg + kSmall;
                        if (nt == 0) {  // a growth without re-growth: its accept log is its region
                            tk = rg;
                            nt = n;
                        }
                        if (lane < nt) tk0 = tk[lane];
                        if (lane < n) rg0 = rg[lane];
                        if (status == kStRect && lane < (int)(sizeof(LsdRect) / 4))
                            rc0 = reinterpret_cast<const unsigned int*>(buf >= 0 ? &my_pool_rect[buf] : &my_small_rect[slot])[lane];
                        bool conflict = false;
                        if (lane < nt) {
                            const unsigned o = (tk0 >> 16) * (unsigned)g.W + (tk0 & 0xffffu);
                            conflict = ((vused[o >> 5] >> (o & 31)) & 1u) != 0;
                        }
                        #pragma unroll 1
                        for (int i = lane + 32; i < nt; i += 32) {
                            const unsigned pp = tk[i];
                            const unsigned o = (pp >> 16) * (unsigned)g.W + (pp & 0xffffu);
                            conflict |= ((vused[o >> 5] >> (o & 31)) & 1u) != 0;
                        }
                        redo = __any_sync(FULL, conflict);
                    }
                    if (redo) {
                        // everything before this ticket is committed: grown now, its growth is the sequential one.  The kernel body
                        // makes the call; this function's state waits in shared memory.
                        if (buf >= 0 && lane == 0) atomicOr(&s_ctl.free_mask, 1ull << buf);
                        __syncwarp();
                        __threadfence_block();
                        if (!kIssue && lane <= r) s_slot[(head + lane) & (kSlots2 - 1)].y = kSlotFree;
                        __syncwarp();
                        __threadfence_block();
                        head += r;
                        if (lane == 0) {
                            ctl->commit_head = head;  // (the claim rule of the grower compares stamps with it)
                            s_view[0].ticket = head;
                            Q.t_next = t_next; Q.head = head; Q.base = base; Q.n_rect = n_rect;
                            Q.n_commit = n_commit; Q.n_void = n_void; Q.n_regrow = n_regrow; Q.n_defer = n_defer;
                            Q.done_mask = done_mask;
                            Q.all_issued = all_issued;
                            Q.c_commit = c_commit + (prof ? clock64() - c0 : 0); Q.c_regrow = c_regrow; Q.c_issue = c_issue; Q.c_idle = c_idle;
                            Q.t_start = t_start;
                            Q.g0 = prof ? clock64() : 0;
                        }
                        __syncwarp();
                        return pix;
                    }
                    if (lane < n) {
                        const unsigned o = (rg0 >> 16) * (unsigned)g.W + (rg0 & 0xffffu);
                        atomicOr(&s_used[o >> 5], 1u << (o & 31));
                    }
                    #pragma unroll 1
                    for (int i = lane + 32; i < n; i += 32) {
                        const unsigned pp = rg[i];
                        const unsigned o = (pp >> 16) * (unsigned)g.W + (pp & 0xffffu);
                        atomicOr(&s_used[o >> 5], 1u << (o & 31));
                    }
                    if (status == kStRect) {
                        if (n_rect < g.seg_cap) {
                            if (lane < (int)(sizeof(LsdRect) / 4)) reinterpret_cast<unsigned int*>(&q[n_rect].rec)[lane] = rc0;
                            n_rect++;
                            if (early_nfa && (n_rect & (kNfaChunk - 1)) == 0) g2_publish_chunk(B, f, n_rect / kNfaChunk - 1);
                        } else if (lane == 0) {
                            atomicOr(B.flags + f, 1);
                        }
                    }
                    n_commit++;
                    if (buf >= 0 && lane == 0) atomicOr(&s_ctl.free_mask, 1ull << buf);
                    __syncwarp();  // the next ticket of the run reads the map this one wrote
                }
                // the slots go back to Free before the window moves on: a committer that is not the issuer must never take the Done of
                // a slot's previous ticket for the state of its next one
                if (!kIssue && lane < run) s_slot[(head + lane) & (kSlots2 - 1)].y = kSlotFree;
                __syncwarp();
                __threadfence_block();
                head += run;
                if (lane == 0) ctl->commit_head = head;
                progress = true;
                if (prof) c_commit += clock64() - c0;
            }
        }
        // ---------------- issue tickets: a burst of chunks of the seed list, published once ----------------
        if (kIssue && !all_issued) {
            const long long i0 = prof ? clock64() : 0;
            int room = min(gs.window - (t_next - head), gs.lookahead - (t_next - ctl->grow_next));
            const int t0 = t_next;
            int chunks = 8;  // then look at the head of the window again
            while (room > 0 && chunks-- > 0) {
                const bool isfree = base + lane < ns && !((done_mask >> lane) & 1u) && !((vused[pa >> 5] >> (pa & 31)) & 1u);
                const unsigned m = __ballot_sync(FULL, isfree);
                const int cnt = __popc(m);
                const int take = min(cnt, room);
                const int r = __popc(m & lt);
                if (isfree && r < take) {
                    volatile int4* sl = &s_slot[(t_next + r) & (kSlots2 - 1)];
                    sl->x = (int)pa;
                    sl->y = kSlotReady;
                }
                t_next += take;
                room -= take;
                if (take < cnt) {  // the window / the look-ahead ends inside this chunk: the rest of it is looked at again later
                    const unsigned sel = __ballot_sync(FULL, isfree && r == take - 1);
                    done_mask |= (2u << (__ffs(sel) - 1)) - 1u;
                    break;
                }
                base += 32;
                done_mask = 0;
                pa = pb; pb = pc; pc = pd;
                pd = base + 96 + lane < ns ? sd[base + 96 + lane] : 0;
                if (base >= ns) {
                    all_issued = true;
                    break;
                }
            }
            if (t_next != t0) {
                __threadfence_block();
                __syncwarp();
                if (lane == 0) ctl->ticket_next = t_next;
            }
            if (t_next != t0 || chunks < 7) progress = true;
            if (prof) c_issue += clock64() - i0;
        }
        if (kIssue) {
            if (all_issued && head == t_next) break;
        } else if (head == t_next && ctl->all_issued) {
            __threadfence_block();
            t_next = ctl->ticket_next;  // the final count: all_issued is written after it
            if (head == t_next) break;
        }
        if (!progress) {
            const long long i0 = prof ? clock64() : 0;
            __nanosleep(gs.poll_ns);
            if (prof) c_idle += clock64() - i0;
        }
    }
    // ---- the frame is finished: publish its rectangles for the tail helpers, release the growers ----
    if (lane == 0) {
        B.n_rects[f] = n_rect;
        if (B.phase_cycles) {
            long long* pc8 = B.phase_cycles + (size_t)f * 16;
            pc8[0] = c_regrow;
            pc8[1] = clock64() - t_start;
            pc8[2] = c_commit - c_regrow;
            pc8[3] = t_next;
            pc8[4] = n_regrow;
            pc8[5] = n_commit;
            pc8[6] = n_defer;
            pc8[7] = n_void;
            pc8[8] = kIssue ? c_issue : (long long)g2s.stat[5];
            pc8[9] = c_idle;
            pc8[10] = (long long)g2s.stat[0];
            pc8[11] = (long long)g2s.stat[1];
            pc8[12] = (long long)g2s.stat[2];
            pc8[13] = (blockDim.x >> 5) - 1 - gs.split;
            pc8[14] = (long long)g2s.stat[3];
            pc8[15] = (long long)g2s.stat[4];
        }
    }
    __threadfence();
    const int nr = min(n_rect, g.seg_cap);
    const int nchunks = (nr + kNfaChunk - 1) / kNfaChunk;
    const int first = early_nfa ? min(nr / kNfaChunk, kNfaChunksPerFrame) : 0;  // (the full chunks are in the queue already)
    if (nchunks > first && nchunks <= kNfaChunksPerFrame) {
        int b0 = 0;
        if (lane == 0) b0 = atomicAdd(B.nfa_ctl + 0, nchunks - first);
        b0 = __shfl_sync(FULL, b0, 0);
        for (int i = first + lane; i < nchunks; i += 32) B.nfa_items[b0 + i - first] = (unsigned)f * kNfaChunksPerFrame + (unsigned)i;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        atomicAdd(B.nfa_ctl + 2, 1);
        ctl->frame_done = 1;
    }
    return -1;
}

// =============================== issuer (warp 1 when the sequencer's two duties are split) ===============================
// Walks the seed list and gives every seed that is unused in the committed map the next ticket, as far as the window and the
// growers' demand allow.  Committing and issuing are both serial per frame but independent of each other: on two warps a frame
// that is alone on its SM (tracking mode) is no longer bound by their sum.
template <bool kProf>
__device__ __noinline__ void g2_issuer(int f) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    const Grow2Smem& gs = g2s.gs;
    volatile int4* s_slot = reinterpret_cast<volatile int4*>(s_raw);
    const volatile unsigned int* vused = reinterpret_cast<unsigned int*>(s_raw + gs.off_used());
    volatile Grow2Ctl* ctl = &g2s.ctl;
    const int ns = g2s.ctl.ns, window = gs.window, lookahead = gs.lookahead, poll_ns = gs.poll_ns;
    const unsigned int* sd = g2s.B.seeds + (size_t)f * g2s.B.plane;
    int t_next = 0, base = 0;
    unsigned done_mask = 0;
    long long c_issue = 0;
    unsigned pa = lane < ns ? sd[lane] : 0, pb = 32 + lane < ns ? sd[32 + lane] : 0, pc = 64 + lane < ns ? sd[64 + lane] : 0,
             pd = 96 + lane < ns ? sd[96 + lane] : 0;
    bool all_issued = ns == 0;
    while (!all_issued) {
        const long long i0 = kProf ? clock64() : 0;
        int room = min(window - (t_next - ctl->commit_head), lookahead - (t_next - ctl->grow_next));
        const int t0 = t_next;
        int chunks = 16;
        while (room > 0 && chunks-- > 0) {
            const bool isfree = base + lane < ns && !((done_mask >> lane) & 1u) && !((vused[pa >> 5] >> (pa & 31)) & 1u);
            const unsigned m = __ballot_sync(FULL, isfree);
            const int cnt = __popc(m);
            const int take = min(cnt, room);
            const int r = __popc(m & lt);
            if (isfree && r < take) {
                volatile int4* sl = &s_slot[(t_next + r) & (kSlots2 - 1)];
                sl->x = (int)pa;
                sl->y = kSlotReady;
            }
            t_next += take;
            room -= take;
            if (take < cnt) {  // the window / the look-ahead ends inside this chunk: the rest of it is looked at again later
                const unsigned sel = __ballot_sync(FULL, isfree && r == take - 1);
                done_mask |= (2u << (__ffs(sel) - 1)) - 1u;
                break;
            }
            base += 32;
            done_mask = 0;
            pa = pb; pb = pc; pc = pd;
            pd = base + 96 + lane < ns ? sd[base + 96 + lane] : 0;
            if (base >= ns) {
                all_issued = true;
                break;
            }
        }
        if (t_next != t0) {
            __threadfence_block();
            __syncwarp();
            if (lane == 0) ctl->ticket_next = t_next;
        }
        if (kProf) c_issue += clock64() - i0;
        if (t_next == t0 && chunks >= 15 && !all_issued) __nanosleep(poll_ns);
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) {
        if (kProf) g2s.stat[5] = (unsigned long long)c_issue;
        ctl->all_issued = 1;
    }
    while (!ctl->frame_done) __nanosleep(1000);  // (the end-of-frame barrier is for every warp)
}

// =============================== grower ===============================
template <bool kProf>
__device__ __noinline__ int g2_grower(int f, int mybuf) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const Grow2Smem& gs = g2s.gs;
    const GrowBufs& B = g2s.B;
    volatile int4* s_slot = reinterpret_cast<volatile int4*>(s_raw);
    const volatile unsigned int* vused = reinterpret_cast<unsigned int*>(s_raw + gs.off_used());
    unsigned int* s_tiny = reinterpret_cast<unsigned int*>(s_raw + gs.off_tiny());
    volatile Grow2Ctl* ctl = &g2s.ctl;
    unsigned char* s_mine = s_raw + gs.off_growers() + (size_t)(warp - 1 - gs.split) * gs.per_grower();
    LsdFrame* view = reinterpret_cast<LsdFrame*>(s_mine + gs.off_view());
    GrowResult* res = reinterpret_cast<GrowResult*>(s_mine + gs.off_res());
    unsigned int* my_pool_reg = B.pool_reg + (size_t)blockIdx.x * gs.pool_n * kSpecCap2;
    unsigned int* my_pool_touched = B.pool_touched + (size_t)blockIdx.x * gs.pool_n * kSpecCap2;
    const LsdPix* rec = B.rec + (size_t)f * B.plane;
    const unsigned int* ring = reinterpret_cast<unsigned int*>(s_mine + gs.off_ring());
    const int min_reg_size = g2s.g.min_reg_size, poll_ns = gs.poll_ns, W = g2s.g.W;
    constexpr bool prof = kProf;  // the clock64 accounting of pl_line_grow_phases is compiled into a copy of its own
    if (lane == 0) {
        view->ang = B.angdeg + (size_t)f * B.plane;
        view->g2 = B.g2 + (size_t)f * B.plane;
        view->rec = B.rec + (size_t)f * B.plane;
        view->cs0 = B.cs0 + (size_t)f * B.plane;
    }
    __syncwarp();
    while (true) {
        while (mybuf < 0) {  // a ticket is only claimed with a buffer in hand (commits free buffers in ticket order)
            mybuf = pool_pop(&g2s.ctl.free_mask, lane);
            if (mybuf < 0) __nanosleep(200);
        }
        int t = 0;
        if (lane == 0) t = atomicAdd(&g2s.ctl.grow_next, 1);
        t = __shfl_sync(FULL, t, 0);
        bool over = false;
        const long long w0 = prof ? clock64() : 0;
        while (true) {
            if (ctl->ticket_next > t) break;
            if (ctl->frame_done) { over = true; break; }
            __nanosleep(poll_ns);
        }
        if (over) break;
        const long long w1 = prof ? clock64() : 0;
        __threadfence_block();
        volatile int4* sl = &s_slot[t & (kSlots2 - 1)];
        const int pix = sl->x;
        // already swallowed, or stamped by an uncommitted earlier ticket (most likely being swallowed): not grown now; the
        // sequencer decides when its turn comes
        const unsigned cl = *(const volatile unsigned int*)&rec[pix].claim & 0xffffu;
        const unsigned d = (unsigned)(t - (int)cl) & 0xffffu;
        if ((((vused[pix >> 5] >> (pix & 31)) & 1u) != 0) || (d != 0 && d <= (unsigned)(t - ctl->commit_head))) {
            if (lane == 0) sl->y = slot_pack(kSlotDone, kStDeferred, -1);
            __syncwarp();
            continue;
        }
        if (lane == 0) {
            view->reg = my_pool_reg + (size_t)mybuf * kSpecCap2;
            view->touched = nullptr;  // the first growth is not logged: its log is the region
            view->touched_buf = my_pool_touched + (size_t)mybuf * kSpecCap2;
            view->ticket = t;
        }
        __syncwarp();
        int r_n, r_nt = 0, r_status = kStNoRect;
        {
            double reg_angle;
            r_n = lsd_region_grow_t<true>(*view, pix % W, pix / W, kPiD * 22.5 / 180, &reg_angle, r_nt);
            if (r_n < 0) r_status = r_n;
            else if (r_n >= min_reg_size) r_status = lsd_fit_refine(*view, &r_n, reg_angle, &r_nt, res);
            lsd_priv_reset(*view, r_nt);
        }
        const long long w2 = prof ? clock64() : 0;
        // a large region keeps the buffer until it is committed; a small one moves to the slot's small buffer, a tiny one
        // (the ring still holds every point of it) into the slot itself
        const bool tiny = r_status == kStNoRect && r_n <= kTiny && r_nt == 0;
        const bool small = !tiny && r_status >= 0 && r_n <= kSmall && r_nt <= kSmall;
        const bool keep = r_status >= 0 && !small && !tiny;
        const int slot = t & (kSlots2 - 1);
        if (tiny) {
            if (lane < r_n) s_tiny[slot * kTiny + lane] = ring[lane];
        } else if (small) {
            const unsigned int* sr = my_pool_reg + (size_t)mybuf * kSpecCap2;  // (reduce_region_radius reorders the list in place)
            const unsigned int* st = my_pool_touched + (size_t)mybuf * kSpecCap2;
            unsigned int* dst = B.small_buf + ((size_t)blockIdx.x * kSlots2 + slot) * 2 * kSmall;
            #pragma unroll 1
            for (int i = lane; i < r_n; i += 32) dst[i] = sr[i];
            #pragma unroll 1
            for (int i = lane; i < r_nt; i += 32) dst[kSmall + i] = st[i];
            if (r_status == kStRect && lane < (int)(sizeof(LsdRect) / 4))
                reinterpret_cast<unsigned int*>(B.small_rect + (size_t)blockIdx.x * kSlots2 + slot)[lane] = reinterpret_cast<const unsigned int*>(&res->rec)[lane];
        } else if (keep && r_status == kStRect) {
            if (lane < (int)(sizeof(LsdRect) / 4))
                reinterpret_cast<unsigned int*>(B.pool_rect + (size_t)blockIdx.x * gs.pool_n + mybuf)[lane] = reinterpret_cast<const unsigned int*>(&res->rec)[lane];
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) {
            sl->z = r_n;
            sl->w = r_nt;
            __threadfence_block();
            sl->y = slot_pack(kSlotDone, r_status, keep ? mybuf : -1);
        }
        __syncwarp();
        if (keep) mybuf = -1;
        if (prof && lane == 0) {
            atomicAdd(&g2s.stat[0], (unsigned long long)(w2 - w1));
            atomicAdd(&g2s.stat[1], (unsigned long long)(w1 - w0));
            atomicAdd(&g2s.stat[2], (unsigned long long)(clock64() - w2));
            if (r_status < 0) atomicAdd(&g2s.stat[3], (unsigned long long)(w2 - w1));
            if (tiny) atomicAdd(&g2s.stat[4], (unsigned long long)(w2 - w1));
        }
    }
    return mybuf;
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_lsd_grow2(LineGeom g_, Grow2Smem gs_, int nf_, GrowBufs B_) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const unsigned FULL = 0xffffffffu;
    if (threadIdx.x == 0) {
        g2s.g = g_;
        g2s.gs = gs_;
        g2s.B = B_;
        g2s.nf = nf_;
        g2s.ctl.free_mask = (gs_.pool_n >= 64 ? ~0ull : ((1ull << gs_.pool_n) - 1ull)) & ~((1ull << (NW - 1 - gs_.split)) - 1ull);
    }
    const int g0 = 1 + gs_.split;  // the first grower warp
    if (warp >= g0) {  // the static part of a grower's frame view, and its empty private bitmap
        unsigned char* s_mine = s_raw + gs_.off_growers() + (size_t)(warp - g0) * gs_.per_grower();
        unsigned char* dir = s_mine + gs_.off_dir();
        #pragma unroll 1
        for (int i = lane; i < gs_.tiles; i += 32) dir[i] = 0xffu;
        if (lane == 0) {
            LsdFrame F;
            F.sval = reinterpret_cast<float2*>(s_mine);
            F.ring = reinterpret_cast<unsigned int*>(s_mine + gs_.off_ring());
            F.scratch = reinterpret_cast<double*>(s_mine + gs_.off_scratch());
            F.pool = reinterpret_cast<unsigned int*>(s_mine + gs_.off_pool());
            F.rev = reinterpret_cast<unsigned short*>(s_mine + gs_.off_rev());
            F.dir = dir;
            F.ntiles = reinterpret_cast<int*>(s_mine + gs_.off_ntiles());
            *F.ntiles = 0;
            F.tw = (g_.W + 31) >> 5;
            F.pool_tiles = gs_.pool_tiles;
            F.sparse = true;
            F.bits = nullptr;
            F.ticket = 0;
            F.W = g_.W;
            F.H = g_.H;
            F.reg = nullptr;
            F.touched = nullptr;
            F.touched_buf = nullptr;
            F.reg_cap = F.touched_cap = kSpecCap2;
            F.used_bits = reinterpret_cast<unsigned int*>(s_raw + gs_.off_used());
            F.commit_head = &g2s.ctl.commit_head;
            F.ang = nullptr; F.g2 = nullptr; F.rec = nullptr; F.cs0 = nullptr;
            *reinterpret_cast<LsdFrame*>(s_mine + gs_.off_view()) = F;
        }
    }
    int mybuf = warp - g0;  // grower i starts with buffer i
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    int* sm_busy = B_.nfa_ctl + 4 + (smid & (kSmBusySlots - 1));
    if (threadIdx.x == 0) atomicAdd(sm_busy, 1);
    while (true) {
        __syncthreads();  // the previous frame is finished, every warp has left it
        if (threadIdx.x == 0) {
            const int fnext = atomicAdd(B_.frame_counter, 1);
            g2s.ctl.frame = fnext;
            g2s.ctl.ns = fnext < nf_ ? B_.n_seeds[fnext] : 0;
            g2s.ctl.ticket_next = g2s.ctl.grow_next = g2s.ctl.commit_head = g2s.ctl.frame_done = g2s.ctl.all_issued = 0;
            for (int k = 0; k < 8; k++) g2s.stat[k] = 0;
        }
        unsigned int* s_used = reinterpret_cast<unsigned int*>(s_raw + gs_.off_used());
        #pragma unroll 1
        for (int i = threadIdx.x; i < gs_.bits_words; i += blockDim.x) s_used[i] = 0;
        if (threadIdx.x < kSlots2) reinterpret_cast<volatile int4*>(s_raw)[threadIdx.x].y = kSlotFree;
        __syncthreads();
        const int f = g2s.ctl.frame;
        if (f >= nf_) break;
        const bool prof = B_.phase_cycles != nullptr;
        if (warp == 0) {
            const bool sp = gs_.split != 0;
            int rq = prof ? (sp ? g2_sequencer<true, false>(f, 0) : g2_sequencer<true, true>(f, 0))
                          : (sp ? g2_sequencer<false, false>(f, 0) : g2_sequencer<false, true>(f, 0));
            while (rq >= 0) {
                lsd_grow_seed(*reinterpret_cast<LsdFrame*>(s_raw + gs_.off_seq() + gs_.off_view()), rq, g_.min_reg_size,
                              reinterpret_cast<GrowResult*>(s_raw + gs_.off_seq() + gs_.off_res()));
                rq = prof ? (sp ? g2_sequencer<true, false>(f, 1) : g2_sequencer<true, true>(f, 1))
                          : (sp ? g2_sequencer<false, false>(f, 1) : g2_sequencer<false, true>(f, 1));
            }
        }
        else if (warp < g0) {
            if (prof) g2_issuer<true>(f);
            else g2_issuer<false>(f);
        } else mybuf = prof ? g2_grower<true>(f, mybuf) : g2_grower<false>(f, mybuf);
    }
    // no frame left for this CTA: validate rectangles of finished frames while other CTAs are still growing.  Once every frame is
    // finished the kernel ends: k_lsd_nfa then validates what is left with the whole GPU.
    if (threadIdx.x == 0) atomicSub(sm_busy, 1);
    if (nf_ > 1 && gs_.tail_nfa) {
        while (true) {
            if (gs_.tail_nfa == 2) {
                // help only from an SM none of whose CTAs still grows: validation next to a grower slows that grower, and the
                // last growers set the kernel's time (300 frames: 36.0 -> 35.4 ms for grow + NFA)
                int st = 0;
                while (true) {
                    if (lane == 0) st = *(volatile int*)(B_.nfa_ctl + 2) >= nf_ ? 2 : (*(volatile int*)sm_busy > 0 ? 1 : 0);
                    st = __shfl_sync(FULL, st, 0);
                    if (st != 1) break;
                    __nanosleep(4000);
                }
                if (st == 2) break;
            }
            int it = 0;
            if (lane == 0) it = atomicAdd(B_.nfa_ctl + 1, 1);
            it = __shfl_sync(FULL, it, 0);
            unsigned item = 0xffffffffu;
            while (true) {  // the item may not be published yet
                if (lane == 0) {
                    if (*(volatile int*)(B_.nfa_ctl + 2) >= nf_) item = 0xfffffffeu;  // every frame finished: stop helping
                    else if (it < *(volatile int*)(B_.nfa_ctl + 0)) item = *(volatile unsigned int*)(B_.nfa_items + it);
                }
                item = __shfl_sync(FULL, item, 0);
                if (item != 0xffffffffu) break;
                __nanosleep(1000);
            }
            if (item == 0xfffffffeu) break;
            __threadfence();
            const bool full = (item >> 31) != 0;
            item &= 0x7fffffffu;
            const int fi = (int)(item / kNfaChunksPerFrame), ch = (int)(item % kNfaChunksPerFrame);
            const int nr = full ? (ch + 1) * kNfaChunk : min(__ldcg(B_.n_rects + fi), g_.seg_cap);  // (written by another SM during this launch: not through L1)
            for (int t = ch * kNfaChunk; t < min(nr, (ch + 1) * kNfaChunk); t++)
                lsd_nfa_one(g_, B_.angdeg + (size_t)fi * B_.plane, B_.queue, B_.qres, B_.qvalid, B_.nfa_tabs, fi, t);
        }
    }
}

}  // namespace pl
