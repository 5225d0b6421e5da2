/*
 * fz_exec_alt.cuh -- round 2's alternative LZ77 executors, all measured slower than k_execute on BASELINE config 2 and none the
 * default (profiles/r02_notes.md section 1): they stay selectable (FZG_EXEC_W = d, s, p2 / p4 / p8) and are covered by the decode
 * tests, because each of them pins down what does NOT bound the stage.
 *
 *   k_execute2          two sequences per lane, half the warps per SM        -> a warp's round is bound by its dependent instruction stream
 *   exec_block_steps    8-byte steps dealt out evenly over the lanes         -> not by lane utilisation or instruction count
 *   k_execute_pass<W>   W warps per frame, one barrier per pass              -> fewer frames in flight do not pay for barriers and lockstep
 *
 * Included by fz_decode.cu inside namespace fz, after the helpers these kernels share with k_execute (ld8_any, st_stage,
 * warp_copy, warp_big_sequence, exec_prefetch, warp_scan_incl, kStage, kExecWarps ...).  Replaces the copy loop of
 * zstd::stream::copy_decode (/root/reference/src/main.rs:463-467) like k_execute does.
 */
#pragma once

// ld8_any in two halves: issue the (one or two) aligned 8-byte loads now, build the value later -- so that several copies' loads
// are in flight before the first of them is used (a warp stalls at the first USE of a load, not at the load)
struct Ld8 { uint2 w0, w1; uint32_t sh; };
__device__ __forceinline__ Ld8 ld8_issue(const uint8_t* g, uint32_t nb)
{
    Ld8 r; r.w0 = make_uint2(0, 0); r.w1 = make_uint2(0, 0); r.sh = 0;
    if (nb) {
        const uintptr_t a = (uintptr_t)g & ~(uintptr_t)7;
        r.sh = (uint32_t)((uintptr_t)g & 7);
        r.w0 = *(const uint2*)a;
        if (r.sh + nb > 8) r.w1 = *(const uint2*)(a + 8);
    }
    return r;
}
__device__ __forceinline__ uint64_t ld8_finish(const Ld8& r) { return funnel8(r.w0.x, r.w0.y, r.w1.x, r.w1.y, r.sh); }

// ---- the same round with TWO sequences per lane (64 per round, sequence k of the round on lane k / 2, slot k % 2) and half the
// warps per SM: the same number of copies in flight per SM from half as many frames, i.e. twice the L2 share per frame.
constexpr uint32_t kStage2 = 2 * kStage;
__device__ __forceinline__ void exec_block_warp2(uint8_t* stage, const Block& b, const uint64_t* __restrict__ sq, uint8_t* g0,
                                                 uint64_t done, int& status, uint32_t lane)
{
    const uint32_t nseq = b.nseq, rsize = b.rsize;
    const uint8_t* __restrict__ lit = b.lit;
    const uint32_t in0 = b.rep_in[0], in1 = b.rep_in[1], in2 = b.rep_in[2];
    uint32_t Ecarry = 0, LEcarry = 0;
    uint64_t rc[2];
#pragma unroll
    for (int q = 0; q < 2; q++) rc[q] = 2 * lane + q < nseq ? __ldg(sq + 2 * lane + q) : 0;
    for (uint32_t g = 0; g < nseq;) {
        const uint32_t nv = min(64u, nseq - g);
        uint32_t E[2], LE[2], S[2], LEp[2], M[2], off[2]; bool valid[2];
#pragma unroll
        for (int q = 0; q < 2; q++) { valid[q] = 2 * lane + q < nv; const uint64_t r = valid[q] ? rc[q] : 0; E[q] = rec_e(r); LE[q] = rec_le(r); off[q] = rec_off(r); }
        const uint32_t ll_ = (nv - 1) >> 1, lq_ = (nv - 1) & 1;        // lane / slot of the round's last sequence
        const uint32_t Elast = __shfl_sync(kFull, lq_ ? E[1] : E[0], ll_), LElast = __shfl_sync(kFull, lq_ ? LE[1] : LE[0], ll_);
#pragma unroll
        for (int q = 0; q < 2; q++) if (!valid[q]) { E[q] = Elast; LE[q] = LElast; }
        S[0] = __shfl_up_sync(kFull, E[1], 1); LEp[0] = __shfl_up_sync(kFull, LE[1], 1);
        if (lane == 0) { S[0] = Ecarry; LEp[0] = LEcarry; }
        S[1] = E[0]; LEp[1] = LE[0];
        const uint32_t gS = Ecarry;                              // output position where this round starts
#pragma unroll
        for (int q = 0; q < 2; q++) {
            M[q] = S[q] + (LE[q] - LEp[q]);
            off[q] = valid[q] ? off_resolve(off[q], in0, in1, in2) : 1;
            if (valid[q] && (uint64_t)off[q] > done + M[q]) { off[q] = 0; status = FZG_E_CORRUPT; }     // reaches before the frame start
        }
        // sequences of this round: the leading ones whose output fits the stage (E never decreases: a prefix in sequence order)
        const uint32_t m = (uint32_t)__popc(__ballot_sync(kFull, valid[0] && E[0] - gS <= kStage2)) + (uint32_t)__popc(__ballot_sync(kFull, valid[1] && E[1] - gS <= kStage2));
        if (m == 0) {                                            // sequence g alone is larger than the stage
            const uint32_t ll0 = __shfl_sync(kFull, LE[0] - LEp[0], 0), ml0 = __shfl_sync(kFull, E[0] - M[0], 0), off0 = __shfl_sync(kFull, off[0], 0);
            warp_big_sequence(g0 + gS, lit + LEcarry, ll0, ml0, off0, lane);
            Ecarry = __shfl_sync(kFull, E[0], 0); LEcarry = __shfl_sync(kFull, LE[0], 0);
            g += 1;
#pragma unroll
            for (int q = 0; q < 2; q++) rc[q] = g + 2 * lane + q < nseq ? __ldg(sq + g + 2 * lane + q) : 0;
            continue;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) rc[q] = g + m + 2 * lane + q < nseq ? __ldg(sq + g + m + 2 * lane + q) : 0;
        const uint32_t el_ = (m - 1) >> 1, eq_ = (m - 1) & 1;
        const uint32_t gE = __shfl_sync(kFull, eq_ ? E[1] : E[0], el_);         // end of the round's output
        const uint32_t LEend = __shfl_sync(kFull, eq_ ? LE[1] : LE[0], el_);
        const uint32_t a = (uint32_t)((uintptr_t)(g0 + gS) & 15); // stage[a + i] <-> g0[gS + i]: same low address bits as HBM
        uint8_t* const st = stage + a - gS;                       // st[p] is the stage byte of output position p (gS <= p < gE)
        bool mine[2];
#pragma unroll
        for (int q = 0; q < 2; q++) mine[q] = 2 * lane + q < m;
        // ---- 1. literal runs
        {
            uint32_t pos[2]; const uint8_t* src[2]; bool go[2];
#pragma unroll
            for (int q = 0; q < 2; q++) { pos[q] = S[q]; src[q] = lit + LEp[q]; go[q] = mine[q] && pos[q] < M[q]; }
            while (__any_sync(kFull, go[0] || go[1])) {
                Ld8 ld[2]; uint32_t nbq[2] = { 0, 0 };
#pragma unroll
                for (int q = 0; q < 2; q++) { nbq[q] = go[q] ? min(8u, M[q] - pos[q]) : 0u; ld[q] = ld8_issue(src[q], nbq[q]); }       // both slots' loads first
#pragma unroll
                for (int q = 0; q < 2; q++)
                    if (go[q]) { st_stage(st + pos[q], ld8_finish(ld[q]), nbq[q]); pos[q] += nbq[q]; src[q] += nbq[q]; go[q] = pos[q] < M[q]; }
            }
        }
        __syncwarp();
        // ---- 2. matches
        {
            uint32_t pos[2]; bool pending[2];
#pragma unroll
            for (int q = 0; q < 2; q++) { pos[q] = M[q]; pending[q] = mine[q] && pos[q] < E[q]; }
            while (__any_sync(kFull, pending[0] || pending[1])) {
                const uint32_t p0 = __ballot_sync(kFull, pending[0]), p1 = __ballot_sync(kFull, pending[1]);
                const uint32_t fl = (uint32_t)__ffs((int)(p0 | p1)) - 1u, fq = (p0 >> fl) & 1u ? 0u : 1u;   // first unfinished sequence, in sequence order
                const uint32_t front = __shfl_sync(kFull, fq ? pos[1] : pos[0], fl);                         // everything below it is written
                bool go[2] = { pending[0], pending[1] };
                while (__any_sync(kFull, go[0] || go[1])) {
                    // both slots' loads are issued before either slot's stores (a slot only reads below `front` or its own bytes)
                    uint64_t vv[2] = { 0, 0 }; uint32_t nbq[2] = { 0, 0 }; Ld8 ld[2]; bool raw[2] = { false, false };
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint8_t* sp = (const uint8_t*)g0; uint32_t nl = 0;        // what to load for this slot (nl == 0: nothing)
                        if (go[q]) {
                            const bool is_first = lane == fl && (uint32_t)q == fq;
                            uint32_t nb = min(8u, E[q] - pos[q]);
                            uint64_t v = 0;
                            const uint32_t of = off[q];
                            if (of == 0) { /* corrupt: zeros */ }
                            else if (of < 8 && of < nb) {         // the step overlaps itself: expand the period byte by byte
                                const int32_t s0 = (int32_t)pos[q] - (int32_t)of;
                                const bool ok = is_first || (uint32_t)(s0 + (int32_t)of) <= front || s0 >= (int32_t)M[q];   // period written?
                                if (ok) {
                                    const uint8_t* pp = s0 < (int32_t)gS ? (const uint8_t*)g0 + s0 : (const uint8_t*)st + s0;
                                    const uint32_t take = s0 < (int32_t)gS ? min(of, gS - (uint32_t)s0) : of;   // a period straddling the round start
                                    uint64_t pat = ld8_any(pp, take);
                                    if (take < of) pat = (pat & ((1ull << (8 * take)) - 1)) | (ld8_any((const uint8_t*)st + gS, of - take) << (8 * take));
                                    for (uint32_t i = 0; i < nb; i++) v |= ((pat >> (8 * (i % of))) & 0xFF) << (8 * i);
                                } else nb = 0;
                            } else {
                                const int32_t s = (int32_t)pos[q] - (int32_t)of;
                                // available bytes: below `front`, or this sequence's own match bytes written so far
                                const uint32_t lim = is_first ? pos[q] : ((s >= (int32_t)M[q]) ? pos[q] : front);
                                if (s < (int32_t)gS) { nb = min(nb, gS - (uint32_t)s); sp = (const uint8_t*)g0 + s; nl = nb; }       // before the round: HBM / L2
                                else if ((uint32_t)s + nb <= lim) { sp = (const uint8_t*)st + s; nl = nb; }
                                else if ((uint32_t)s < lim) { nb = lim - (uint32_t)s; sp = (const uint8_t*)st + s; nl = nb; }
                                else nb = 0;
                            }
                            vv[q] = v; nbq[q] = nb;
                        }
                        raw[q] = nl != 0;
                        ld[q] = ld8_issue(sp, nl);                 // both slots' loads are issued before either is used
                    }
#pragma unroll
                    for (int q = 0; q < 2; q++) if (raw[q]) vv[q] = ld8_finish(ld[q]);
#pragma unroll
                    for (int q = 0; q < 2; q++)
                        if (go[q]) {
                            if (nbq[q]) { st_stage(st + pos[q], vv[q], nbq[q]); pos[q] += nbq[q]; go[q] = pos[q] < E[q]; }
                            else go[q] = false;                   // its source is still being produced by a lower sequence
                        }
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 2; q++) pending[q] = mine[q] && pos[q] < E[q];
            }
        }
        __syncwarp();
        // ---- 3. flush stage[a .. a + (gE - gS)) -> g0 + gS: head bytes, aligned 16-byte body, tail bytes
        {
            const uint32_t n = gE - gS;
            uint8_t* gd = g0 + gS;
            const uint32_t head = min(n, (16 - a) & 15);
            if (lane < head) gd[lane] = stage[a + lane];
            const uint32_t nvec = (n - head) >> 4;
            for (uint32_t i = lane; i < nvec; i += 32) *(uint4*)(gd + head + 16 * i) = *(const uint4*)(stage + a + head + 16 * i);
            const uint32_t tail0 = head + (nvec << 4);
            if (tail0 + lane < n) gd[tail0 + lane] = stage[a + tail0 + lane];
        }
        __syncwarp();
        Ecarry = gE; LEcarry = LEend;
        g += m;
#if FZ_EXEC_PREFETCH
        {   // the next round's match sources, a round ahead (see exec_block_warp)
            const uint32_t nn = g < nseq ? min(64u, nseq - g) : 0u;
            const uint32_t En0 = rec_e(rc[0]), LEn0 = rec_le(rc[0]), En1 = rec_e(rc[1]), LEn1 = rec_le(rc[1]);
            uint32_t Sn0 = __shfl_up_sync(kFull, En1, 1), LEpn0 = __shfl_up_sync(kFull, LEn1, 1);
            if (lane == 0) { Sn0 = Ecarry; LEpn0 = LEcarry; }
            const uint32_t Mn0 = Sn0 + (LEn0 - LEpn0), Mn1 = En0 + (LEn1 - LEn0);
            const uint32_t o0 = off_resolve(rec_off(rc[0]), in0, in1, in2), o1 = off_resolve(rec_off(rc[1]), in0, in1, in2);
            if (2 * lane < nn && o0 != 0 && (uint64_t)o0 <= done + Mn0) exec_prefetch(g0 + Mn0 - o0, En0 - Mn0);
            if (2 * lane + 1 < nn && o1 != 0 && (uint64_t)o1 <= done + Mn1) exec_prefetch(g0 + Mn1 - o1, En1 - Mn1);
        }
#endif
    }
    // literals after the last sequence
    warp_copy(g0 + Ecarry, lit + LEcarry, rsize - Ecarry, lane);
    __syncwarp();
}

__global__ void __launch_bounds__(kExecWarps * 32, kExecCtasPerSm / 2) k_execute2(Frame* frames, const Block* blocks, const Item* items,
                                                                                  const ItemOut* outs, const uint64_t* seqs,
                                                                                  uint32_t n_frames, uint32_t* ticket)
{
    __shared__ __align__(16) uint8_t s_stage[kExecWarps][kStage2 + 48];
    const uint32_t lane = threadIdx.x & 31;
    uint8_t* stage = s_stage[threadIdx.x >> 5];
    for (;;) {
        uint32_t f = 0;
        if (lane == 0) f = atomicAdd(ticket, 1);
        f = __shfl_sync(kFull, f, 0);
        if (f >= n_frames) return;
        Frame& fr = frames[f];
        if (outs[fr.item].fail) continue;
        uint8_t* const fbase = items[fr.item].dst + fr.out_off;
        uint64_t done = 0;
        int status = 0;
        for (uint32_t kb = 0; kb < fr.n_blocks; kb++) {
            const Block& b = blocks[fr.first_block + kb];
            uint8_t* const g0 = fbase + done;
            const uint32_t rsize = b.rsize;
            if (b.type == BT_RAW) warp_copy(g0, b.src, rsize, lane);
            else if (b.type == BT_RLE) {
                const uint8_t v = b.src[0];
                for (uint32_t i = lane; i < rsize; i += 32) g0[i] = v;
            } else if (b.nseq == 0) warp_copy(g0, b.lit, rsize, lane);
            else exec_block_warp2(stage, b, seqs + b.seq_base, g0, done, status, lane);
            __syncwarp();                      // later blocks read this one back (the window)
            done += rsize;
        }
        status = __reduce_max_sync(kFull, status);
        if (lane == 0 && status) fr.status = status;
    }
}

// ---- the same round, with the 8-byte STEPS of its copies dealt out evenly over the lanes.
// exec_block_warp gives a lane one sequence and lets it loop over its literal run and its match: a round then takes as many
// step iterations as its longest run plus its longest match (about eight on JSON text, 21 of 32 lanes busy) for an average
// of 1.7 steps per sequence.  Here a scan over the step counts of the round's sequences numbers the steps (position order),
// lane l takes steps l and l + 32 (a round holds at most 64), finds their sequence with a five-shuffle binary search over
// the scan and fetches its fields with three shuffles: two iterations for the same round.  A step is atomic (<= 8 bytes, one
// source); a match step that would read its own match (offset < length) takes its bytes from the last period BEFORE the
// match instead (x[p] = x[M - off + (p - M) mod off]), so a step only ever depends on bytes below its sequence's match.
// Steps whose source lies inside the round wait for the frontier (the first unfinished step, in position order) to pass it.
__device__ __forceinline__ void exec_block_steps(uint8_t* stage, const Block& b, const uint64_t* __restrict__ sq, uint8_t* g0,
                                                 uint64_t done, int& status, uint32_t lane)
{
    const uint32_t nseq = b.nseq, rsize = b.rsize;
    const uint8_t* __restrict__ lit = b.lit;
    const uint32_t in0 = b.rep_in[0], in1 = b.rep_in[1], in2 = b.rep_in[2];
    uint32_t Ecarry = 0, LEcarry = 0;
    uint64_t rcur = lane < nseq ? __ldg(sq + lane) : 0;          // records of the current round; the next round's are loaded a round early
    for (uint32_t g = 0; g < nseq;) {
        const uint32_t nv = min(32u, nseq - g);
        const uint64_t r = lane < nv ? rcur : 0;
        uint32_t E = rec_e(r), LE = rec_le(r);
        const uint32_t Elast = __shfl_sync(kFull, E, nv - 1), LElast = __shfl_sync(kFull, LE, nv - 1);
        if (lane >= nv) { E = Elast; LE = LElast; }
        uint32_t S = __shfl_up_sync(kFull, E, 1), LEp = __shfl_up_sync(kFull, LE, 1);
        if (lane == 0) { S = Ecarry; LEp = LEcarry; }
        const uint32_t gS = Ecarry;                              // output position where this round starts
        const uint32_t ll = LE - LEp, M = S + ll, ml = E - M;
        uint32_t off = lane < nv ? off_resolve(rec_off(r), in0, in1, in2) : 1;
        if (lane < nv && (uint64_t)off > done + M) { off = 0; status = FZG_E_CORRUPT; }     // reaches before the frame start
        // steps of every sequence, numbered in position order
        const uint32_t c = lane < nv ? ((ll + 7) >> 3) + ((ml + 7) >> 3) : 0u;
        const uint32_t cs = warp_scan_incl(c, lane);
        // sequences of this round: the leading ones whose output fits the stage and whose steps fit two per lane
        const uint32_t fit = __ballot_sync(kFull, lane < nv && E - gS <= kStage && cs <= 64u);
        const uint32_t m = fit == kFull ? 32u : (uint32_t)__ffs((int)~fit) - 1u;
        if (m == 0) {                                            // sequence g alone is larger than the stage
            const uint32_t ll0 = __shfl_sync(kFull, ll, 0), ml0 = __shfl_sync(kFull, ml, 0), off0 = __shfl_sync(kFull, off, 0);
            warp_big_sequence(g0 + gS, lit + LEcarry, ll0, ml0, off0, lane);
            Ecarry = __shfl_sync(kFull, E, 0); LEcarry = __shfl_sync(kFull, LE, 0);
            g += 1;
            rcur = g + lane < nseq ? __ldg(sq + g + lane) : 0;
            continue;
        }
        rcur = g + m + lane < nseq ? __ldg(sq + g + m + lane) : 0;
        const uint32_t gE = __shfl_sync(kFull, E, m - 1);         // end of the round's output
        const uint32_t n_steps = __shfl_sync(kFull, cs, m - 1);
        const uint32_t a = (uint32_t)((uintptr_t)(g0 + gS) & 15); // stage[a + i] <-> g0[gS + i]: same low address bits as HBM
        uint8_t* const st = stage + a - gS;                       // st[p] is the stage byte of output position p (gS <= p < gE)
        // a sequence in three words: (S - gS) [0:10) | ll [10:20) | ml [20:30);  distance;  LEp [0:18) | first step [18:25)
        const uint32_t w1 = (S - gS) | (ll << 10) | (ml << 20), w3 = LEp | ((cs - c) << 18);
        // ---- a lane's two steps: pos / n = destination, src = literal offset or match distance, kind: 0 none, 1 literal, 2 match
        uint32_t posA = 0, nA = 0, srcA = 0, mA = 0, kindA = 0, posB = 0, nB = 0, srcB = 0, mB = 0, kindB = 0;
        auto deal = [&](uint32_t k, uint32_t& pos, uint32_t& n, uint32_t& src, uint32_t& mstart, uint32_t& kind) {
            uint32_t j = 0;                                       // the sequence of step k: the first one whose inclusive count exceeds k
#pragma unroll
            for (uint32_t stp = 16; stp; stp >>= 1) { const uint32_t t = __shfl_sync(kFull, cs, j + stp - 1); if (t <= k) j += stp; }
            const uint32_t v1 = __shfl_sync(kFull, w1, j), v2 = __shfl_sync(kFull, off, j), v3 = __shfl_sync(kFull, w3, j);
            const uint32_t oS = gS + (v1 & 1023u), oll = (v1 >> 10) & 1023u, oml = v1 >> 20, oM = oS + oll;
            const uint32_t t = k - (v3 >> 18), nl = (oll + 7) >> 3;
            kind = 0;
            if (k < n_steps) {
                if (t < nl) { kind = 1; pos = oS + 8 * t; n = min(8u, oM - pos); src = (v3 & 0x3FFFFu) + 8 * t; }
                else { kind = 2; pos = oM + 8 * (t - nl); n = min(8u, oM + oml - pos); src = v2; mstart = oM; }
            }
        };
        deal(lane, posA, nA, srcA, mA, kindA);
        if (n_steps > 32) deal(lane + 32, posB, nB, srcB, mB, kindB);
        // ---- one attempt at a step; returns true when it is stored.  front: every byte below it is written (stage or HBM)
        auto attempt = [&](uint32_t pos, uint32_t n, uint32_t src, uint32_t mstart, uint32_t kind, uint32_t front) -> bool {
            uint64_t v = 0;
            if (kind == 1) v = ld8_any(lit + src, n);
            else {
                const uint32_t off_ = src;
                if (off_ == 0) { /* corrupt: zeros */ }
                else if (pos - mstart + n > off_) {               // the step would read its own match: the last period before it
                    if (mstart > front) return false;
                    uint32_t ph = (pos - mstart) % off_;
                    for (uint32_t i = 0; i < n; i++) {
                        const int32_t q = (int32_t)mstart - (int32_t)off_ + (int32_t)ph;
                        const uint8_t by = q < (int32_t)gS ? *((const uint8_t*)g0 + q) : *((const uint8_t*)st + q);
                        v |= (uint64_t)by << (8 * i);
                        if (++ph == off_) ph = 0;
                    }
                } else {
                    const int32_t s = (int32_t)pos - (int32_t)off_;
                    if (s + (int32_t)n <= (int32_t)gS) v = ld8_any((const uint8_t*)g0 + s, n);             // before the round: HBM / L2
                    else if ((uint32_t)(s + (int32_t)n) > front) return false;                                // not written yet
                    else if (s >= (int32_t)gS) v = ld8_any((const uint8_t*)st + s, n);
                    else {                                            // straddles the round start
                        const uint32_t n0 = gS - (uint32_t)s;
                        v = (ld8_any((const uint8_t*)g0 + s, n0) & ((1ull << (8 * n0)) - 1ull)) | (ld8_any((const uint8_t*)st + gS, n - n0) << (8 * n0));
                    }
                }
            }
            st_stage(st + pos, v, n);
            return true;
        };
        bool pendA = kindA != 0, pendB = kindB != 0;
        uint32_t front = gS;
        for (;;) {
            if (pendA) pendA = !attempt(posA, nA, srcA, mA, kindA, front);
            if (__any_sync(kFull, pendB)) { if (pendB) pendB = !attempt(posB, nB, srcB, mB, kindB, front); }
            __syncwarp();
            const uint32_t pa = __ballot_sync(kFull, pendA), pb = __ballot_sync(kFull, pendB);
            if (!(pa | pb)) break;
            // the frontier: where the first unfinished step (position order: slot A lanes 0..31, then slot B) begins
            const uint32_t fl = pa ? (uint32_t)__ffs((int)pa) - 1u : (uint32_t)__ffs((int)pb) - 1u;
            front = __shfl_sync(kFull, pa ? posA : posB, fl);
        }
        // ---- flush stage[a .. a + (gE - gS)) -> g0 + gS: head bytes, aligned 16-byte body, tail bytes
        {
            const uint32_t n = gE - gS;
            uint8_t* gd = g0 + gS;
            const uint32_t head = min(n, (16 - a) & 15);
            if (lane < head) gd[lane] = stage[a + lane];
            const uint32_t nvec = (n - head) >> 4;
            for (uint32_t i = lane; i < nvec; i += 32) *(uint4*)(gd + head + 16 * i) = *(const uint4*)(stage + a + head + 16 * i);
            const uint32_t tail0 = head + (nvec << 4);
            if (tail0 + lane < n) gd[tail0 + lane] = stage[a + tail0 + lane];
        }
        __syncwarp();
        Ecarry = gE; LEcarry = __shfl_sync(kFull, LE, m - 1);
        g += m;
#if FZ_EXEC_PREFETCH
        {   // the next round's match sources: asked for now, a whole round before they are read (see exec_block_warp)
            const uint32_t nn = g < nseq ? min(32u, nseq - g) : 0u;
            const uint32_t En = rec_e(rcur), LEn = rec_le(rcur);
            uint32_t Sn = __shfl_up_sync(kFull, En, 1), LEpn = __shfl_up_sync(kFull, LEn, 1);
            if (lane == 0) { Sn = Ecarry; LEpn = LEcarry; }
            const uint32_t Mn = Sn + (LEn - LEpn);
            const uint32_t offn = off_resolve(rec_off(rcur), in0, in1, in2);
            if (lane < nn && offn != 0 && (uint64_t)offn <= done + Mn) exec_prefetch(g0 + Mn - offn, En - Mn);
        }
#endif
    }
    // literals after the last sequence
    warp_copy(g0 + Ecarry, lit + LEcarry, rsize - Ecarry, lane);
    __syncwarp();
}


// ------------------------------------------------------------------ execute, a few warps per frame, barrier passes
// k_execute keeps 32 frames in flight per SM and is bound by the DRAM fetches of their window reads (15 % L2 hit rate, see
// DESIGN.md section 2).  Here W (2 or 4) warps share a frame, so 32 / W frames are in flight per SM at the same number of warps,
// WITHOUT the per-byte bookkeeping of k_execute_cta: the round is k_execute's round with 32 W sequences, one per thread, and its
// frontier loop with the warp replaced by the CTA -- every pass ends in one barrier at which each warp posts its first
// unfinished thread; the minimum is the CTA's first unfinished thread and the position it has reached is the frontier of the
// next pass.  Sources before the round come from HBM / L2 at once, so nearly everything is done in the first pass.
template <int W> struct ExecPass {
    static constexpr uint32_t T = W * 32, stage = W * kStage, stage_bytes = stage + 48;
    static constexpr int ctas_per_sm = 32 / W;
};

template <int W>
__device__ __forceinline__ void exec_block_pass(uint8_t* stage, uint32_t* cnt, uint32_t* fr /*[2][W]*/, const Block& b, const uint64_t* __restrict__ sq,
                                                uint8_t* g0, uint64_t done, int& status, uint32_t tid)
{
    constexpr uint32_t T = ExecPass<W>::T, kCtaStage = ExecPass<W>::stage;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t nseq = b.nseq, rsize = b.rsize;
    const uint8_t* __restrict__ lit = b.lit;
    const uint32_t in0 = b.rep_in[0], in1 = b.rep_in[1], in2 = b.rep_in[2];
    uint32_t Ecarry = 0, LEcarry = 0, fpar = 0;                  // CTA-uniform
    uint64_t rcur = tid < nseq ? __ldg(sq + tid) : 0;            // the round's records; the next round's are loaded a round early
    uint64_t rprev = (lane == 0 && warp > 0 && tid <= nseq) ? __ldg(sq + tid - 1) : 0;   // record before a warp's first one
    for (uint32_t g = 0; g < nseq;) {
        const uint32_t nv = min(T, nseq - g);
        const uint64_t rl = __ldg(sq + g + nv - 1);
        const uint32_t Elast = rec_e(rl), LElast = rec_le(rl);
        const uint64_t r = tid < nv ? rcur : 0;
        uint32_t E = rec_e(r), LE = rec_le(r);
        if (tid >= nv) { E = Elast; LE = LElast; }
        uint32_t S = __shfl_up_sync(kFull, E, 1), LEp = __shfl_up_sync(kFull, LE, 1);
        if (lane == 0) {
            if (warp == 0) { S = Ecarry; LEp = LEcarry; }
            else if (tid <= nv) { S = rec_e(rprev); LEp = rec_le(rprev); }
            else { S = Elast; LEp = LElast; }
        }
        const uint32_t gS = Ecarry;                              // output position where this round starts
        uint32_t off = tid < nv ? off_resolve(rec_off(r), in0, in1, in2) : 1;
        if (tid < nv && (uint64_t)off > done + S + (LE - LEp)) { off = 0; status = FZG_E_CORRUPT; }   // reaches before the frame start
        // sequences of this round: the leading ones whose output fits the stage (E never decreases)
        uint32_t n = nv;
        if (Elast - gS > kCtaStage) {
            const uint32_t fit = __ballot_sync(kFull, tid < nv && E - gS <= kCtaStage);
            if (lane == 0) cnt[warp] = (uint32_t)__popc(fit);
            __syncthreads();
            n = 0;
            for (int w = 0; w < W; w++) n += cnt[w];
            __syncthreads();
        }
        if (n == 0) {                                            // sequence g alone is larger than the stage
            const uint64_t r0 = __ldg(sq + g);
            const uint32_t E0 = rec_e(r0), LE0 = rec_le(r0), M0 = gS + (LE0 - LEcarry);
            uint32_t off0 = off_resolve(rec_off(r0), in0, in1, in2);
            if ((uint64_t)off0 > done + M0) off0 = 0;
            group_copy(g0 + gS, lit + LEcarry, LE0 - LEcarry, tid, T);
            __syncthreads();
            {
                uint8_t* m = g0 + M0; const uint32_t ml0 = E0 - M0;
                if (off0 != 0) {
                    const uint8_t* sp = m - off0;
                    if (off0 >= 16 * T) {                                // source and destination of a 16 T-byte round never overlap
                        for (uint32_t i = 0; i < ml0; i += 16 * T) {
                            const uint32_t nb = min(16u, ml0 > i + 16 * tid ? ml0 - i - 16 * tid : 0u);
                            for (uint32_t k = 0; k < nb; k++) m[i + 16 * tid + k] = sp[i + 16 * tid + k];
                            __syncthreads();
                        }
                    } else for (uint32_t i = tid; i < ml0; i += T) m[i] = sp[i % off0];   // periodic; the period lies below m: written
                }
            }
            __syncthreads();
            Ecarry = E0; LEcarry = LE0;
            g += 1;
            rcur = g + tid < nseq ? __ldg(sq + g + tid) : 0;
            rprev = (lane == 0 && warp > 0 && g + tid <= nseq) ? __ldg(sq + g + tid - 1) : 0;
            continue;
        }
        const uint64_t re = __ldg(sq + g + n - 1);
        const uint32_t gE = rec_e(re), LEend = rec_le(re);       // end of the round's output / literals
        rcur = g + n + tid < nseq ? __ldg(sq + g + n + tid) : 0;
        rprev = (lane == 0 && warp > 0 && g + n + tid <= nseq) ? __ldg(sq + g + n + tid - 1) : 0;
        const bool mine = tid < n;
        if (!mine) { S = gE; E = gE; LE = LEend; LEp = LEend; }
        const uint32_t M = S + (LE - LEp);
        const uint32_t a = (uint32_t)((uintptr_t)(g0 + gS) & 15); // stage[a + i] <-> g0[gS + i]: same low address bits as HBM
        uint8_t* const st = stage + a - gS;                       // st[p] is the stage byte of output position p (gS <= p < gE)
        // ---- 1. literal runs
        {
            uint32_t pos = S; const uint8_t* src = lit + LEp;
            bool go = mine && pos < M;
            while (__any_sync(kFull, go)) {
                if (go) {
                    const uint32_t nb = min(8u, M - pos);
                    st_stage(st + pos, ld8_any(src, nb), nb);
                    pos += nb; src += nb; go = pos < M;
                }
            }
        }
        // ---- 2. matches, in passes
        {
            uint32_t pos = M;
            bool pending = mine && pos < E;
            for (;;) {
                // the CTA's first unfinished thread and the position it has reached (key = tid << 16 | pos - gS)
                const uint32_t pm = __ballot_sync(kFull, pending);
                const uint32_t wl = pm ? (uint32_t)__ffs((int)pm) - 1u : 0u;
                const uint32_t wpos = __shfl_sync(kFull, pos, wl);
                if (lane == 0) fr[fpar * W + warp] = pm ? (((warp << 5) | wl) << 16) | (wpos - gS) : 0xFFFFFFFFu;
                __syncthreads();                                  // (also: the stage bytes of the previous pass / the literal runs are visible)
                uint32_t key = 0xFFFFFFFFu;
#pragma unroll
                for (int w = 0; w < W; w++) key = min(key, fr[fpar * W + w]);
                fpar ^= 1u;
                if (key == 0xFFFFFFFFu) break;
                const uint32_t first = key >> 16, front = gS + (key & 0xFFFFu);   // every output byte below `front` is written (HBM or stage)
                bool go = pending;
                while (__any_sync(kFull, go)) {
                    if (go) {
                        uint32_t nb = min(8u, E - pos);
                        uint64_t v = 0;
                        if (off == 0) { /* corrupt: zeros */ }
                        else if (off < 8 && off < nb) {           // the step overlaps itself: expand the period byte by byte
                            const int32_t s0 = (int32_t)pos - (int32_t)off;
                            const bool ok = tid == first || (uint32_t)(s0 + (int32_t)off) <= front || s0 >= (int32_t)M;   // period written?
                            if (ok) {
                                const uint8_t* sp = s0 < (int32_t)gS ? (const uint8_t*)g0 + s0 : (const uint8_t*)st + s0;
                                const uint32_t take = s0 < (int32_t)gS ? min(off, gS - (uint32_t)s0) : off;   // a period straddling the round start
                                uint64_t pat = ld8_any(sp, take);
                                if (take < off) pat = (pat & ((1ull << (8 * take)) - 1)) | (ld8_any((const uint8_t*)st + gS, off - take) << (8 * take));
                                for (uint32_t i = 0; i < nb; i++) v |= ((pat >> (8 * (i % off))) & 0xFF) << (8 * i);
                            } else nb = 0;
                        } else {
                            const int32_t s = (int32_t)pos - (int32_t)off;
                            // available bytes: below `front`, or this thread's own match bytes written so far
                            const uint32_t lim = tid == first ? pos : ((s >= (int32_t)M) ? pos : front);
                            if (s < (int32_t)gS) {                 // before the round: HBM / L2 (earlier rounds, earlier blocks)
                                nb = min(nb, gS - (uint32_t)s);    // a step straddling the round start is split
                                v = ld8_any((const uint8_t*)g0 + s, nb);
                            } else if ((uint32_t)s + nb <= lim) v = ld8_any((const uint8_t*)st + s, nb);
                            else if ((uint32_t)s < lim) { nb = lim - (uint32_t)s; v = ld8_any((const uint8_t*)st + s, nb); }
                            else nb = 0;
                        }
                        if (nb) { st_stage(st + pos, v, nb); pos += nb; go = pos < E; }
                        else go = false;                          // its source is still being produced by a lower thread
                    }
                }
                pending = mine && pos < E;
            }
        }
        // ---- 3. flush stage[a .. a + (gE - gS)) -> g0 + gS: head bytes, aligned 16-byte body, tail bytes (the break above came after a barrier)
        {
            const uint32_t nby = gE - gS;
            uint8_t* gd = g0 + gS;
            const uint32_t head = min(nby, (16 - a) & 15);
            if (tid < head) gd[tid] = stage[a + tid];
            const uint32_t nvec = (nby - head) >> 4;
            for (uint32_t i = tid; i < nvec; i += T) *(uint4*)(gd + head + 16 * i) = *(const uint4*)(stage + a + head + 16 * i);
            const uint32_t tail0 = head + (nvec << 4);
            if (tail0 + tid < nby) gd[tail0 + tid] = stage[a + tail0 + tid];
        }
        __syncthreads();
        Ecarry = gE; LEcarry = LEend;
        g += n;
#if FZ_EXEC_PREFETCH
        {   // the next round's match sources: asked for now (see k_execute)
            const uint32_t nn = g < nseq ? min(T, nseq - g) : 0u;
            const uint32_t En = rec_e(rcur), LEn = rec_le(rcur);
            uint32_t Sn = __shfl_up_sync(kFull, En, 1), LEpn = __shfl_up_sync(kFull, LEn, 1);
            if (lane == 0) { if (warp == 0) { Sn = Ecarry; LEpn = LEcarry; } else { Sn = rec_e(rprev); LEpn = rec_le(rprev); } }
            const uint32_t Mn = Sn + (LEn - LEpn);
            const uint32_t offn = off_resolve(rec_off(rcur), in0, in1, in2);
            if (tid < nn && offn != 0 && (uint64_t)offn <= done + Mn && offn > Mn - Ecarry) exec_prefetch(g0 + Mn - offn, En - Mn);
        }
#endif
    }
    // literals after the last sequence
    group_copy(g0 + Ecarry, lit + LEcarry, rsize - Ecarry, tid, T);
}

template <int W>
__global__ void __launch_bounds__(ExecPass<W>::T, ExecPass<W>::ctas_per_sm) k_execute_pass(Frame* frames, const Block* blocks, const Item* items,
                                                                                         const ItemOut* outs, const uint64_t* seqs,
                                                                                         uint32_t n_frames, uint32_t* ticket)
{
    constexpr uint32_t T = ExecPass<W>::T;
    __shared__ __align__(16) uint8_t s_stage[ExecPass<W>::stage_bytes];
    __shared__ uint32_t s_cnt[W], s_fr[2 * W], s_f;
    const uint32_t tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_f = atomicAdd(ticket, 1);
        __syncthreads();
        const uint32_t f = s_f;
        if (f >= n_frames) return;
        Frame& fr = frames[f];
        if (outs[fr.item].fail) continue;
        uint8_t* const fbase = items[fr.item].dst + fr.out_off;
        uint64_t done = 0;
        int status = 0;
        for (uint32_t kb = 0; kb < fr.n_blocks; kb++) {
            const Block& b = blocks[fr.first_block + kb];
            uint8_t* const g0 = fbase + done;
            const uint32_t rsize = b.rsize;
            if (b.type == BT_RAW) group_copy(g0, b.src, rsize, tid, T);
            else if (b.type == BT_RLE) {
                const uint8_t v = b.src[0];
                for (uint32_t i = tid; i < rsize; i += T) g0[i] = v;
            } else if (b.nseq == 0) group_copy(g0, b.lit, rsize, tid, T);
            else exec_block_pass<W>(s_stage, s_cnt, s_fr, b, seqs + b.seq_base, g0, done, status, tid);
            __syncthreads();                               // later blocks read this one back (the window)
            done += rsize;
        }
        if (__syncthreads_or(status) && tid == 0) fr.status = FZG_E_CORRUPT;
    }
}

