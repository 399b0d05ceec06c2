// inflate_host.cpp -- the device inflater's core (zotmer_b200/csrc/inflate_core.cuh), compiled for the host with a
// "warp" of ONE lane, so that its bit reader, table construction, header parsing and copy rules can be checked
// against zlib without a GPU (tests/test_host_layer.py).  Test infrastructure only: nothing in the product loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../zotmer_b200/csrc/inflate_core.cuh"

extern "C" int zi_inflate_host(const uint8_t* src, uint32_t clen, uint8_t* out, uint32_t isize) {
    // the device reads whole words around the stream: give the copy the same slack, at an odd alignment
    uint8_t* buf = (uint8_t*)calloc(clen + 64, 1);
    const int skew = (int)(clen % 4);
    memcpy(buf + 8 + skew, src, clen);
    zinf::Scratch* S = (zinf::Scratch*)calloc(1, sizeof(zinf::Scratch));
    const int rc = zinf::inflate_member<1>(0, 1u, buf + 8 + skew, clen, out, isize, S);
    free(S);
    free(buf);
    return rc;
}
