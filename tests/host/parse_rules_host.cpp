// Host-side check of the FASTA line-state rules (zotmer_b200/csrc/fasta_rules.cuh) -- no GPU needed.
// Usage: parse_rules_host <file.fa> <out.codes>
// Emulates what fasta_kernel does with scans, but sequentially over 16-byte pieces, and verifies
// that the scanned transfer maps reproduce the byte-exact sequential state.  Writes the dense code
// stream (0..3 bases, 4 break) that the kernel would emit.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../zotmer_b200/csrc/fasta_rules.cuh"
using namespace zb;

static uint4 load16(const std::vector<uint8_t>& d, size_t off) {
    uint8_t b[16];
    for (int i = 0; i < 16; i++) b[i] = (off + i < d.size()) ? d[off + i] : 0x20;
    uint4 v;
    memcpy(&v, b, 16);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> d;
    uint8_t buf[65536];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + r);
    fclose(f);
    const size_t n = d.size(), np = (n + 15) / 16;
    std::vector<FaMasks> mk(np);
    std::vector<uint32_t> sum(np);
    for (size_t i = 0; i < np; i++) { mk[i] = fa_masks(load16(d, i * 16)); sum[i] = fa_summary(mk[i]); }
    // associativity / apply consistency of the transfer maps, exhaustively over 7-bit maps that occur
    for (size_t i = 0; i + 2 < np; i++) {
        uint32_t a = sum[i], b = sum[i + 1], c = sum[i + 2];
        if (fa_compose(fa_compose(a, b), c) != fa_compose(a, fa_compose(b, c))) { fprintf(stderr, "compose not associative at %zu\n", i); return 1; }
        for (uint32_t st = 0; st < 8; st++)
            if (fa_apply(fa_compose(a, b), st) != fa_apply(b, fa_apply(a, st))) { fprintf(stderr, "apply/compose mismatch at %zu\n", i); return 1; }
    }
    // backward carries, exact (sequential from the right); beyond the end = 1
    std::vector<uint8_t> bin(np);
    {
        uint32_t b = 1;
        for (size_t i = np; i-- > 0;) {
            bin[i] = (uint8_t)b;
            FaPiece p0 = fa_piece(mk[i], false, false, false);
            uint32_t gb = p0.B & 1u, pb = (sum[i] & FA_PF) ? 1u : 0u;
            uint32_t viaSummary = gb | (pb & b);
            FaPiece pe = fa_piece(mk[i], false, false, b);
            if ((pe.B & 1u) != viaSummary) { fprintf(stderr, "backward summary mismatch at piece %zu\n", i); return 1; }
            b = viaSummary;
        }
    }
    std::vector<uint8_t> out;
    uint32_t st = 1u;         // exact sequential state (f=1 at start)
    uint32_t comp = FA_IDENT;  // composition of all summaries so far
    unsigned long long nrec = 0;
    for (size_t i = 0; i < np; i++) {
        if (fa_apply(comp, 1u) != st) { fprintf(stderr, "scan state mismatch at piece %zu: %u vs %u\n", i, fa_apply(comp, 1u), st); return 1; }
        FaPiece p = fa_piece(mk[i], st & 1u, st & 2u, bin[i]);
        uint32_t live = (st & 4u) ? 0xffffu : (p.HS ? (0xffffu & ~((p.HS & (0u - p.HS)) - 1u)) : 0u);
        uint32_t skip = mk[i].W & (p.F | p.B);
        uint32_t emit = live & (p.HS | (~p.HIN & ~skip & 0xffffu));
        uint4 v = load16(d, i * 16);
        for (int b = 0; b < 16; b++) {
            if (i * 16 + b >= n) break;
            if ((p.HS >> b) & 1u) nrec++;
            if ((emit >> b) & 1u) out.push_back(((p.HS >> b) & 1u) ? 4 : (uint8_t)code_of(byte_of(v, b)));
        }
        uint32_t nst = 0;
        if (p.F & 0x8000u) nst |= 1u;
        if (p.HIN & 0x8000u) nst |= 2u;
        if ((st & 4u) || p.HS) nst |= 4u;
        st = nst;
        comp = fa_compose(comp, sum[i]);
    }
    FILE* o = fopen(argv[2], "wb");
    fwrite(out.data(), 1, out.size(), o);
    fclose(o);
    printf("%llu\n", nrec);
    return 0;
}
