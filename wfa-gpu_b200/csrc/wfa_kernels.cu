/*
 * wfa_kernels.cu -- hand-written sm_100a kernels of the gap-affine WFA hot path.
 * See wfa_kernels.cuh for the kernel list and the reference lines each replaces.
 */
#include "wfa_kernels.cuh"

namespace wfagpu {

/* ======================================================================== */
/*                               PTX helpers                                */
/* ======================================================================== */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

/* 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP). */
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/* ======================================================================== */
/*                               pack kernel                                */
/* ======================================================================== */
/*
 * One warp per sequence.  Per iteration a warp consumes 512 ASCII bytes: lane i
 * reads two 16-byte aligned uint4 (its own and the next one), picks the 24
 * bytes that start at sequence byte 16*i (the sequence start is only 4-byte
 * aligned, lib/aligner.c:127-166), and emits two packed words:
 *   word j = bases [8j, 8j+16), base 8j in bits 31:30, code (c & 6) >> 1
 * (A=0 C=1 T=2 G=3, lib/kernels/sequence_packing_kernel.cu:79).  The 8-base
 * stride makes every extend start with >= 9 bases in a single word.
 * The has_N flag follows sequence_packing_kernel.cu:54-76 literally.
 */
__device__ __forceinline__ uint32_t pack4(uint32_t w)
{
    /* 4 ASCII bytes (first base in the low byte) -> 8 bits, first base in bits 7:6 */
    return (((w >> 1) & 0x03030303u) * 0x40100401u) >> 24;
}

__device__ __forceinline__ uint32_t keep_bytes(uint32_t w, int nbytes)
{
    /* keep the first nbytes (low) bytes of a little-endian word, zero the rest */
    if (nbytes >= 4) return w;
    if (nbytes <= 0) return 0u;
    return w & (0xffffffffu >> (32 - 8 * nbytes));
}

__global__ void __launch_bounds__(256) pack_kernel(PackParams p)
{
    const int lane = threadIdx.x & 31;
    const uint32_t seq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seq >= 2u * p.n_pairs) return;
    wfagpu_pair_t *pr = p.pairs + (seq >> 1);
    const bool is_text = seq & 1u;
    const uint32_t len = is_text ? pr->tlen : pr->plen;
    const uint32_t a_off = is_text ? pr->t_ascii : pr->p_ascii;
    uint32_t *dst = p.packed + (is_text ? pr->t_word : pr->p_word);
    const uint32_t nwords = ((len + 7u) >> 3) + 1u;          /* words that carry data or the terminator */
    const uint32_t nwords_pad = (nwords + 3u) & ~3u;           /* TMA copies whole 16-byte units */
    const char *src = p.ascii + a_off;
    const uintptr_t addr = (uintptr_t)src;
    const uint4 *base16 = (const uint4 *)(addr & ~(uintptr_t)15);
    const int q = (int)((addr & 15) >> 2);                      /* misalignment in words: 0..3 */
    /* groups of 4 bytes the reference's flag loop visits */
    const uint32_t ngroups = (len + (4u - (len & 3u))) >> 2;
    /* number of aligned uint4 that may be touched without leaving the batch buffer */
    const uint32_t n16 = (uint32_t)(((addr & 15) + len + 1 + 15) >> 4);
    bool flag = (len >= (1u << 15));

    for (uint32_t it = 0; it * 64u < nwords_pad; ++it) {
        const uint32_t u = it * 32u + lane;                     /* index of this lane's first uint4 */
        uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0), c = make_uint4(0, 0, 0, 0);
        if (u < n16) a = __ldg(base16 + u);
        if (u + 1 < n16) b = __ldg(base16 + u + 1);
        if (q == 3 && u + 2 < n16) c = __ldg(base16 + u + 2);
        uint32_t w[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x};
        uint32_t s[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            uint32_t v = w[j];
            v = (q == 1) ? w[j + 1] : v;
            v = (q == 2) ? w[j + 2] : v;
            v = (q == 3) ? w[j + 3] : v;
            s[j] = v;
        }
        /* s[j] = sequence bytes [16u*... ] i.e. bytes 16*u + 4j .. +3 */
        const uint32_t byte0 = 16u * u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t g = (byte0 >> 2) + j;                /* 4-byte group index */
            if (g < ngroups && s[j] != 0u) {
                const uint32_t t = s[j] ^ 0x4e4e4e4eu;
                const uint32_t f = (t & 0xff) & ((t >> 8) & 0xff) & ((t >> 16) & 0xff) & (t >> 24);
                flag |= (f == 0u);
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int nb = (int)len - (int)(byte0 + 4u * j);
            s[j] = keep_bytes(s[j], nb);
        }
        const uint32_t w0 = (pack4(s[0]) << 24) | (pack4(s[1]) << 16) | (pack4(s[2]) << 8) | pack4(s[3]);
        const uint32_t w1 = (pack4(s[2]) << 24) | (pack4(s[3]) << 16) | (pack4(s[4]) << 8) | pack4(s[5]);
        const uint32_t j0 = 2u * u;
        if (j0 + 1 < nwords_pad) {
            *reinterpret_cast<uint2 *>(dst + j0) = make_uint2(w0, w1);
        } else if (j0 < nwords_pad) {
            dst[j0] = w0;
        }
    }
    flag = __any_sync(0xffffffffu, flag);
    if (flag && lane == 0) atomicOr(&pr->flags, WFAGPU_PAIR_HAS_N);
}

void launch_pack(const PackParams &p, cudaStream_t s)
{
    if (p.n_pairs == 0) return;
    const uint32_t warps = 2u * p.n_pairs;
    const uint32_t blocks = (warps + 7u) / 8u;
    pack_kernel<<<blocks, 256, 0, s>>>(p);
}

/* ======================================================================== */
/*                            alignment kernel                              */
/* ======================================================================== */

template <bool WARP>
struct Group {
    __device__ static __forceinline__ int tid() { return WARP ? (threadIdx.x & 31) : threadIdx.x; }
    __device__ static __forceinline__ int size() { return WARP ? 32 : blockDim.x; }
    __device__ static __forceinline__ int id_in_cta() { return WARP ? (threadIdx.x >> 5) : 0; }
    __device__ static __forceinline__ void sync()
    {
        if (WARP) __syncwarp(); else __syncthreads();
    }
};

/* Bounded common prefix on packed words (replaces WF_extend_kernel,
 * lib/kernels/common_alignment_kernels.cuh:29-111): XOR + count-leading-zeros
 * on 32-bit windows; a window starts inside a single word thanks to the
 * 8-base stride layout, so the common case costs two shared loads. */
__device__ __forceinline__ int extend_packed(const uint32_t *__restrict__ P, const uint32_t *__restrict__ T,
                                             int plen, int tlen, int k, int off)
{
    int v = off - k, h = off;
    const int rem = min(plen - v, tlen - h);
    if (rem < 0) return kOffNull;
    int acc = 0;
    while (true) {
        const int vs = v & 7, hs = h & 7;
        const uint32_t wp = P[v >> 3] << (2 * vs);
        const uint32_t wt = T[h >> 3] << (2 * hs);
        const int nvalid = 16 - max(vs, hs);
        int eq = __clz((int)(wp ^ wt)) >> 1;
        eq = min(eq, nvalid);
        acc += eq;
        if (eq < nvalid || acc >= rem) break;
        v += eq;
        h += eq;
    }
    return off + min(acc, rem);
}

/* Byte-compare variant for pairs the packer flagged (non-ACGT bytes): plain
 * byte equality like the CPU WFA the reference falls back to
 * (utils/wfa_cpu.c:57-85), straight from the ASCII copy in global memory. */
__device__ __forceinline__ int extend_ascii(const char *__restrict__ P, const char *__restrict__ T, int plen,
                                            int tlen, int k, int off)
{
    int v = off - k, h = off;
    const int rem = min(plen - v, tlen - h);
    if (rem < 0) return kOffNull;
    int acc = 0;
    while (acc < rem && P[v + acc] == T[h + acc]) ++acc;
    return off + acc;
}

template <bool ASCII>
__device__ __forceinline__ int extend_any(const void *P, const void *T, int plen, int tlen, int k, int off)
{
    if (ASCII) return extend_ascii((const char *)P, (const char *)T, plen, tlen, k, off);
    return extend_packed((const uint32_t *)P, (const uint32_t *)T, plen, tlen, k, off);
}

struct GroupCtl {
    uint64_t bar[2];     /* TMA completion barriers, one per sequence stage */
    uint32_t idx[2];     /* pair index staged in each buffer                */
    uint32_t bytes_p[2]; /* unused by consumers; kept for debugging         */
    uint32_t n_ops;
    uint32_t ops_off;
};

template <bool WARP, bool ASCII>
__global__ void __launch_bounds__(WARP ? 256 : 1024, 1) wfa_exact_kernel(const __grid_constant__ KernelParams p)
{
    using G = Group<WARP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = G::tid();
    const int gsz = G::size();
    const int lane = threadIdx.x & 31;
    const int groups_per_cta = WARP ? (blockDim.x >> 5) : 1;
    const uint32_t group = blockIdx.x * groups_per_cta + G::id_in_cta();

    /* ---- carve shared memory: [rings][seq stage 0][seq stage 1][ctl] per group ---- */
    const int rows = p.A + 2 * p.E1;
    const size_t ring_bytes = ((size_t)rows * p.row_stride * sizeof(int16_t) + 15) & ~(size_t)15;
    const size_t seq_bytes = (size_t)p.seq_words * 4;          /* one sequence, one stage */
    const size_t group_bytes = ring_bytes + (ASCII ? 0 : 2 * p.stages * seq_bytes) + sizeof(GroupCtl);
    unsigned char *gbase = smem_raw + (size_t)G::id_in_cta() * ((group_bytes + 15) & ~(size_t)15);
    int16_t *ring = reinterpret_cast<int16_t *>(gbase);
    uint32_t *seqbuf = reinterpret_cast<uint32_t *>(gbase + ring_bytes);
    GroupCtl *ctl = reinterpret_cast<GroupCtl *>(gbase + ring_bytes + (ASCII ? 0 : 2 * p.stages * seq_bytes));

    int16_t *const Mring = ring + p.center;
    int16_t *const Iring = Mring + (size_t)p.A * p.row_stride;
    int16_t *const Dring = Iring + (size_t)p.E1 * p.row_stride;

    uint4 *const arena = p.arena + (size_t)group * p.arena_units;
    uint32_t *const scratch = p.ops_scratch + (size_t)group * p.ops_scratch_words;

    const int x = p.x, o = p.o, e = p.e, A = p.A, E1 = p.E1, GW = p.G;
    const int oe = o + e;

    /* ---- stage-0 prologue: pop the first pair and start its TMA load ---- */
    auto issue_load = [&](int stage, uint32_t idx) {
        /* leader only */
        if (ASCII) return;
        const wfagpu_pair_t pr = p.pairs[idx];
        const uint32_t pw = ((((pr.plen + 7u) >> 3) + 1u) + 3u) & ~3u;
        const uint32_t tw = ((((pr.tlen + 7u) >> 3) + 1u) + 3u) & ~3u;
        uint32_t *dp = seqbuf + (size_t)(2 * stage) * p.seq_words;
        uint32_t *dt = dp + p.seq_words;
        fence_proxy_async();
        mbar_expect_tx(&ctl->bar[stage], (pw + tw) * 4u);
        tma_load_1d(dp, p.packed + pr.p_word, pw * 4u, &ctl->bar[stage]);
        tma_load_1d(dt, p.packed + pr.t_word, tw * 4u, &ctl->bar[stage]);
    };
    auto pop = [&]() -> uint32_t {
        const uint32_t pos = atomicAdd(p.queue, 1u);
        return pos < p.n_items ? p.order[pos] : kInvalidIdx;
    };

    if (tid == 0) {
        if (!ASCII) {
            mbar_init(&ctl->bar[0], 1);
            mbar_init(&ctl->bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        const uint32_t first = pop();
        ctl->idx[0] = first;
        if (first != kInvalidIdx) issue_load(0, first);
    }
    G::sync();

    int stage = 0;
    uint32_t phase_bits = 0; /* bit s = parity to wait for on stage s */

    while (true) {
        const uint32_t idx = ctl->idx[stage];
        if (idx == kInvalidIdx) break;
        if (tid == 0 && p.stages == 2) {
            /* prefetch the next pair into the other stage while this one computes */
            const uint32_t nxt = pop();
            ctl->idx[stage ^ 1] = nxt;
            if (nxt != kInvalidIdx) issue_load(stage ^ 1, nxt);
        }
        const wfagpu_pair_t pr = p.pairs[idx];
        const int plen = (int)pr.plen, tlen = (int)pr.tlen;
        const int kt = tlen - plen;

        const void *Pseq, *Tseq;
        if (ASCII) {
            Pseq = p.ascii + pr.p_ascii;
            Tseq = p.ascii + pr.t_ascii;
        } else {
            Pseq = seqbuf + (size_t)(2 * stage) * p.seq_words;
            Tseq = seqbuf + (size_t)(2 * stage + 1) * p.seq_words;
        }

        /* pairs flagged by the packer are left to the byte-compare launch */
        const bool skip = !ASCII && (pr.flags & WFAGPU_PAIR_HAS_N);

        /* ---- ring prologue: NULL over [-2G, 2G] on every row (no full re-init) ---- */
        {
            const int span = 4 * GW + 1;
            const int total = rows * span;
            for (int i = tid; i < total; i += gsz) {
                const int r = i / span;
                const int k = i - r * span - 2 * GW;
                ring[(size_t)r * p.row_stride + p.center + k] = (int16_t)kOffNull;
            }
        }
        if (!ASCII) mbar_wait(&ctl->bar[stage], (phase_bits >> stage) & 1u);
        phase_bits ^= (1u << stage);
        G::sync();

        int dist = 0;
        bool finished = false;
        unsigned long long my_cells = 0;

        if (!skip) {
            if (tid == 0) Mring[0] = (int16_t)extend_any<ASCII>(Pseq, Tseq, plen, tlen, 0, 0);
            G::sync();
            if (kt == 0 && Mring[0] == tlen) {
                finished = true;
            } else {
                wfagpu_step_t st_next = p.steps[1 < p.d_end ? 1 : 0];
                for (int d = 1; d < p.d_end; ++d) {
                    const wfagpu_step_t st = st_next;
                    if (d + 1 < p.d_end) st_next = p.steps[d + 1];
                    const int n = st.n;
                    if (n > p.n_cap) break;
                    int16_t *const Mc = Mring + (size_t)(d % A) * p.row_stride;
                    int16_t *const Ic = Iring + (size_t)(d % E1) * p.row_stride;
                    int16_t *const Dc = Dring + (size_t)(d % E1) * p.row_stride;

                    if (st.kind == WFAGPU_STEP_NULL) {
                        for (int k = -n - GW + tid; k <= n + GW; k += gsz) {
                            Mc[k] = (int16_t)kOffNull;
                            Ic[k] = (int16_t)kOffNull;
                            Dc[k] = (int16_t)kOffNull;
                        }
                        G::sync();
                        continue;
                    }
                    if (st.kind == WFAGPU_STEP_M) {
                        const int16_t *const Mx = Mring + (size_t)((d - x) % A) * p.row_stride;
                        for (int k = -n - GW + tid; k <= n + GW; k += gsz) {
                            Ic[k] = (int16_t)kOffNull;
                            Dc[k] = (int16_t)kOffNull;
                            int m = kOffNull;
                            if (k >= -n && k <= n) {
                                m = (int)Mx[k] + 1;
                                if (m >= 0) m = extend_any<ASCII>(Pseq, Tseq, plen, tlen, k, m);
                                ++my_cells;
                            }
                            Mc[k] = (int16_t)m;
                        }
                    } else {
                        const int16_t *const Mo = Mring + (size_t)(((d - oe) % A + A) % A) * p.row_stride;
                        const int16_t *const Mx = Mring + (size_t)(((d - x) % A + A) % A) * p.row_stride;
                        const int16_t *const Ie = Iring + (size_t)(((d - e) % E1 + E1) % E1) * p.row_stride;
                        const int16_t *const De = Dring + (size_t)(((d - e) % E1 + E1) % E1) * p.row_stride;
                        uint4 *const row = arena + st.row_off;
                        /* guard cells: NULL on both sides of [-n, n] */
                        for (int g = tid; g < 2 * GW; g += gsz) {
                            const int k = (g < GW) ? (-n - 1 - g) : (n + 1 + (g - GW));
                            Mc[k] = (int16_t)kOffNull;
                            Ic[k] = (int16_t)kOffNull;
                            Dc[k] = (int16_t)kOffNull;
                        }
                        const int width = 2 * n + 1;
                        for (int base = (tid & ~31); base < width; base += gsz) {
                            const int idc = base + lane;
                            const bool in = idc < width;
                            uint32_t bI = 0, bD = 0, bM = 0;
                            if (in) {
                                const int k = idc - n;
                                const int io = (int)Mo[k - 1] + 1;
                                const int ie = (int)Ie[k - 1] + 1;
                                const int pI = max(io * 2, ie * 2 + 1);
                                const int I = pI >> 1;
                                const int dopen = (int)Mo[k + 1];
                                const int dext = (int)De[k + 1];
                                const int pD = max(dopen * 2, dext * 2 + 1);
                                const int D = pD >> 1;
                                const int X = (int)Mx[k] + 1;
                                const int pM = max(max(X * 4 + 2, D * 4 + 3), I * 4 + 1);
                                int M = pM >> 2;
                                if (M >= 0) M = extend_any<ASCII>(Pseq, Tseq, plen, tlen, k, M);
                                Ic[k] = (int16_t)I;
                                Dc[k] = (int16_t)D;
                                Mc[k] = (int16_t)M;
                                bI = pI & 1;
                                bD = pD & 1;
                                bM = pM & 3;
                                ++my_cells;
                            }
                            if (p.with_bt) {
                                const uint32_t m0 = __ballot_sync(0xffffffffu, bI);
                                const uint32_t m1 = __ballot_sync(0xffffffffu, bD);
                                const uint32_t m2 = __ballot_sync(0xffffffffu, bM & 1u);
                                const uint32_t m3 = __ballot_sync(0xffffffffu, bM & 2u);
                                if (lane == 0) row[base >> 5] = make_uint4(m0, m1, m2, m3);
                            }
                        }
                    }
                    G::sync();
                    if (kt >= -n && kt <= n && Mc[kt] == tlen) {
                        finished = true;
                        dist = d;
                        break;
                    }
                }
            }
        }

        /* ---- traceback (leader): decision planes -> 2-bit ops, newest first ---- */
        if (tid == 0) {
            uint32_t n_ops = 0, ops_off = 0;
            if (finished && p.with_bt && dist > 0) {
                int cd = dist, ck = kt, comp = 0;
                uint32_t word = 0;
                __threadfence_block();
                while (!(comp == 0 && cd == 0)) {
                    const wfagpu_step_t st = p.steps[cd];
                    uint32_t op;
                    if (comp == 0) {
                        op = OP_SUB;
                        if (st.kind == WFAGPU_STEP_M) {
                            cd -= x;
                        } else {
                            const int ii = ck + (int)st.n;
                            if (ii < 0 || ii > 2 * (int)st.n) { n_ops = 0; finished = false; break; }
                            const uint4 dec = arena[st.row_off + (ii >> 5)];
                            const int b = ii & 31;
                            const int mop = (int)((dec.z >> b) & 1u) | (int)(((dec.w >> b) & 1u) << 1);
                            if (mop == OP_SUB) cd -= x;
                            else if (mop == OP_INS) comp = 1;
                            else comp = 2;
                        }
                    } else {
                        const int ii = ck + (int)st.n;
                        if (ii < 0 || ii > 2 * (int)st.n || st.kind != WFAGPU_STEP_MDI) { n_ops = 0; finished = false; break; }
                        const uint4 dec = arena[st.row_off + (ii >> 5)];
                        const int b = ii & 31;
                        if (comp == 1) {
                            op = OP_INS;
                            ck -= 1;
                            if ((dec.x >> b) & 1u) cd -= e; else { cd -= oe; comp = 0; }
                        } else {
                            op = OP_DEL;
                            ck += 1;
                            if ((dec.y >> b) & 1u) cd -= e; else { cd -= oe; comp = 0; }
                        }
                    }
                    word |= op << (2 * (n_ops & 15u));
                    ++n_ops;
                    if ((n_ops & 15u) == 0) {
                        scratch[(n_ops >> 4) - 1] = word;
                        word = 0;
                    }
                    if (cd < 0 || (n_ops >> 4) >= p.ops_scratch_words) { n_ops = 0; finished = false; break; }
                }
                if (n_ops & 15u) scratch[n_ops >> 4] = word;
                const uint32_t nw = (n_ops + 15u) >> 4;
                ops_off = atomicAdd(p.ops_pool_head, nw);
                if (ops_off + nw > p.ops_pool_words) { n_ops = 0; finished = false; }
            }
            ctl->n_ops = n_ops;
            ctl->ops_off = ops_off;
            wfagpu_pair_out_t r;
            r.distance = finished ? dist : 0;
            r.ops_off = ops_off;
            r.n_ops = n_ops;
            if (skip) {
                r.status = WFAGPU_ST_NEEDS_ASCII;
                p.ascii_list[atomicAdd(p.ascii_count, 1u)] = idx;
            } else if (finished) {
                r.status = WFAGPU_ST_FINISHED;
            } else {
                r.status = WFAGPU_ST_OVERBUDGET;
                p.retry_list[atomicAdd(p.retry_count, 1u)] = idx;
            }
            p.out[idx] = r;
        }
        if (p.cells) {
            /* optional work counter (profiling builds of the plan only) */
            for (int s = 16; s > 0; s >>= 1) my_cells += __shfl_xor_sync(0xffffffffu, my_cells, s);
            if (lane == 0 && my_cells) atomicAdd(p.cells, my_cells);
        }
        G::sync();
        {
            /* copy the op words from the group's scratch into the pool (coalesced) */
            const uint32_t nw = (ctl->n_ops + 15u) >> 4;
            const uint32_t off = ctl->ops_off;
            for (uint32_t i = tid; i < nw; i += gsz) p.ops_pool[off + i] = scratch[i];
        }
        G::sync();
        if (p.stages == 2) {
            stage ^= 1;
        } else if (tid == 0) {
            /* single buffer (shared memory is tight): fetch the next pair now */
            const uint32_t nxt = pop();
            ctl->idx[0] = nxt;
            if (nxt != kInvalidIdx) issue_load(0, nxt);
        }
        if (p.stages != 2) G::sync();
    }
}

template <bool WARP, bool ASCII>
static cudaError_t launch_one(const KernelParams &p, int threads, int ctas, size_t smem, cudaStream_t s)
{
    auto kfn = wfa_exact_kernel<WARP, ASCII>;
    cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kfn<<<ctas, threads, smem, s>>>(p);
    return cudaGetLastError();
}

size_t exact_smem_bytes(int A, int E1, int row_stride, int seq_words, int groups_per_cta,
                        int stages)
{
    const int rows = A + 2 * E1;
    const size_t ring_bytes = ((size_t)rows * row_stride * sizeof(int16_t) + 15) & ~(size_t)15;
    const size_t seq_bytes = (size_t)seq_words * 4;
    const size_t group_bytes = (ring_bytes + 2 * (size_t)stages * seq_bytes + sizeof(GroupCtl) + 15) & ~(size_t)15;
    return group_bytes * (size_t)groups_per_cta;
}

cudaError_t launch_exact(const KernelParams &p, int group_threads, int groups_per_cta, int ctas,
                         size_t smem_bytes, bool ascii_extend, cudaStream_t s)
{
    const bool warp = (group_threads == 32);
    const int threads = warp ? 32 * groups_per_cta : group_threads;
    if (warp) {
        return ascii_extend ? launch_one<true, true>(p, threads, ctas, smem_bytes, s)
                            : launch_one<true, false>(p, threads, ctas, smem_bytes, s);
    }
    return ascii_extend ? launch_one<false, true>(p, threads, ctas, smem_bytes, s)
                        : launch_one<false, false>(p, threads, ctas, smem_bytes, s);
}

int exact_max_ctas_per_sm(int group_threads, int groups_per_cta, size_t smem_bytes, bool ascii_extend)
{
    const bool warp = (group_threads == 32);
    const int threads = warp ? 32 * groups_per_cta : group_threads;
    int n = 0;
    cudaError_t err;
    if (warp) {
        if (ascii_extend) {
            cudaFuncSetAttribute(wfa_exact_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wfa_exact_kernel<true, true>, threads, smem_bytes);
        } else {
            cudaFuncSetAttribute(wfa_exact_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wfa_exact_kernel<true, false>, threads, smem_bytes);
        }
    } else {
        if (ascii_extend) {
            cudaFuncSetAttribute(wfa_exact_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wfa_exact_kernel<false, true>, threads, smem_bytes);
        } else {
            cudaFuncSetAttribute(wfa_exact_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wfa_exact_kernel<false, false>, threads, smem_bytes);
        }
    }
    if (err != cudaSuccess) return 0;
    return n;
}

} // namespace wfagpu
