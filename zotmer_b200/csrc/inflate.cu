// inflate.cu -- block-compressed input (.gz written by bgzip: BGZF) is inflated ON THE DEVICE.
//
// Replaces zotmer/library/file.py:93-97 (`gunzip -c` child + pipe feeding readFastq / readFasta) for BGZF files:
// the COMPRESSED bytes cross PCIe (a quarter of the text), one warp inflates one member (inflate_core.cuh), and the
// text lands where the parser reads it.  Also here: the record-aligned cut of a text buffer that lives on the device
// (what library/reads.py:pieces does on the host for plain files), so that a file larger than one piece is fed group
// of members by group of members with the incomplete last record carried over.
#include "inflate_core.cuh"
#include "kernels.h"

namespace zb {

static constexpr int INF_THREADS = 256;

// W lanes inflate one member: a whole warp (W = 32), or an aligned half / quarter of one (W = 16 / 8: the groups of a
// warp then run their serial decodes in the same issue slots wherever they take the same branch -- ZB_INFLATE_W)
template <int W>
__global__ void __launch_bounds__(INF_THREADS)
bgzf_inflate_kernel(const uint8_t* __restrict__ comp, const BgzfMember* __restrict__ tab, uint32_t members, uint8_t* out,
                    unsigned int* __restrict__ err /*[0] failures, [1] 1 + first failing member, [2] its code*/) {
    extern __shared__ __align__(16) unsigned char inf_raw[];
    zinf::Scratch* scratch = reinterpret_cast<zinf::Scratch*>(inf_raw);   // [INF_THREADS / W], 6.5 KB each
    constexpr int GROUPS = INF_THREADS / W;
    const uint32_t grp = threadIdx.x / W;
    const uint32_t m = blockIdx.x * GROUPS + grp;
    if (m >= members) return;
    const BgzfMember t = tab[m];
    if (t.isize == 0 && t.clen <= 2) return;   // the empty member that ends a BGZF file
    const int lane = (int)(threadIdx.x % W);
    const uint32_t smask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << ((threadIdx.x & 31u) / W * W));
    const int rc = zinf::inflate_member<W>(lane, smask, comp + t.src, t.clen, out + t.dst, t.isize, &scratch[grp]);
    if (rc != zinf::ZI_OK && lane == 0) {
        if (atomicAdd(&err[0], 1u) == 0) {
            err[1] = m + 1;
            err[2] = (unsigned)rc;
        }
    }
}

template <int W>
static void launch_inflate(Ctx* c, const uint8_t* d_comp, const BgzfMember* d_tab, uint32_t members, uint8_t* d_out, unsigned int* d_err) {
    constexpr int GROUPS = INF_THREADS / W;
    const int smem = (int)(GROUPS * sizeof(zinf::Scratch));   // W = 32: 52 KB, four CTAs = 32 warps = 32 members per SM
    ZB_CUDA(cudaFuncSetAttribute(bgzf_inflate_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    bgzf_inflate_kernel<W><<<(unsigned)div_up(members, GROUPS), INF_THREADS, smem, c->stream>>>(d_comp, d_tab, members, d_out, d_err);
    ZB_LAUNCH_CHECK(c);
}

void bgzf_inflate(Ctx* c, const uint8_t* d_comp, const BgzfMember* d_tab, uint32_t members, uint8_t* d_out, unsigned int* d_err) {
    if (members == 0) return;
    static const int w = [] { const char* e = getenv("ZB_INFLATE_W"); return e ? atoi(e) : 32; }();
    Stage st(c, "inflate");
    if (w == 16) launch_inflate<16>(c, d_comp, d_tab, members, d_out, d_err);
    else if (w == 8) launch_inflate<8>(c, d_comp, d_tab, members, d_out, d_err);
    else launch_inflate<32>(c, d_comp, d_tab, members, d_out, d_err);
}

// ------------------------------------------------------------------------------- record-aligned cut
// res[0] = number of '\n'; res[1] = 1 + the largest i with text[i] == '\n' and text[i + 1] == '>' (0: none)
__global__ void __launch_bounds__(256)
text_scan_kernel(const uint8_t* __restrict__ text, uint64_t n, unsigned long long* __restrict__ res) {
    unsigned long long nl = 0, last = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n; i += stride) {
        if (i + 17 <= n) {
            const uint4 v = *reinterpret_cast<const uint4*>(text + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint8_t nxt = text[i + 16];
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const uint8_t ch = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
                const uint8_t ch1 = (b < 15) ? (uint8_t)(w[(b + 1) >> 2] >> (8 * ((b + 1) & 3))) : nxt;
                if (ch == '\n') {
                    nl++;
                    if (ch1 == '>') last = i + b + 1;
                }
            }
        } else {
            for (uint64_t j = i; j < n; j++) {
                if (text[j] == '\n') {
                    nl++;
                    if (j + 1 < n && text[j + 1] == '>') last = j + 1;
                }
            }
        }
    }
    nl = warp_sum(nl);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long x = __shfl_xor_sync(0xffffffffu, last, o);
        last = x > last ? x : last;
    }
    if (lane_id() == 0) {
        if (nl) atomicAdd(&res[0], nl);
        if (last) atomicMax(&res[1], last);
    }
}

// res[2] = 1 + position of the q-th '\n' counted from the END of the text (0: there are fewer); one warp walks back
__global__ void __launch_bounds__(32)
text_back_kernel(const uint8_t* __restrict__ text, uint64_t n, uint32_t q, unsigned long long* __restrict__ res) {
    const unsigned lane = threadIdx.x;
    uint64_t end = n;
    uint32_t seen = 0;
    while (end > 0) {
        const uint64_t base = end >= 32 ? end - 32 : 0;
        const uint64_t i = base + lane;
        const bool is_nl = (i < end) && text[i] == '\n';
        unsigned bal = __ballot_sync(0xffffffffu, is_nl);
        const uint32_t c = (uint32_t)__popc(bal);
        if (seen + c >= q) {
            // the (q - seen)-th newline from the top of this window
            uint32_t need = q - seen;
            while (need > 1) {
                bal &= ~(1u << (31 - __clz(bal)));
                need--;
            }
            if (lane == 0) res[2] = base + (uint64_t)(31 - __clz(bal)) + 1;
            return;
        }
        seen += c;
        end = base;
    }
    if (lane == 0) res[2] = 0;
}

// Where a text buffer on the device can be cut so that [0, cut) parses like a whole file and [cut, n) is the beginning of
// the next piece.  FASTQ: after the last newline whose number is a multiple of 4; FASTA: in front of the last header line.
// 0 = no cut inside the buffer (one record fills it).  Synchronises.
uint64_t text_cut(Ctx* c, const uint8_t* d_text, uint64_t n, bool is_fasta) {
    if (n == 0) return 0;
    DBuf<unsigned long long> res(c, 4);
    ZB_CUDA(dev_memset(c, res.get(), 0, 32));
    {
        Stage st(c, "text_cut");
        int blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * 8, div_up(n, 256 * 16));
        if (blocks < 1) blocks = 1;
        text_scan_kernel<<<blocks, 256, 0, c->stream>>>(d_text, n, res.get());
        ZB_LAUNCH_CHECK(c);
    }
    ZB_CUDA(read_back(c, res.get(), 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t nl = c->h_scalars[0], hdr = c->h_scalars[1];
    if (is_fasta) return hdr;   // text[hdr] is the '>' of the last header line that follows a newline
    const uint64_t whole = nl & ~3ull;
    if (whole == 0) return 0;
    text_back_kernel<<<1, 32, 0, c->stream>>>(d_text, n, (uint32_t)(nl - whole) + 1u, res.get());
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, res.get() + 2, 8));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    return c->h_scalars[0];
}

}  // namespace zb
