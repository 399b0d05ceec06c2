// inflate_host.cpp -- the device inflater's core (zotmer_b200/csrc/inflate_core.cuh), compiled for the host, so that its
// bit reader, table construction, header parsing and copy rules can be checked against zlib without a GPU
// (tests/test_host_layer.py).  zi_inflate_host: a "warp" of ONE lane.  zi_inflate_host_lanes: W = 2 / 4 / 8 lanes, one
// host thread each, ZI_SYNC a real barrier -- the lanes' protocol (literals parked in the lanes' registers, two-literal
// table entries, shared match copies, synchronisation only where a match reads unsynchronised bytes) runs as it does
// on the device.  Test infrastructure only: nothing in the product loads it.
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static pthread_barrier_t* g_bar = nullptr;
static inline void zi_host_sync() {
    if (g_bar) pthread_barrier_wait(g_bar);
}
#define ZI_HOST_SYNC zi_host_sync
#include "../../zotmer_b200/csrc/inflate_core.cuh"

extern "C" int zi_inflate_host(const uint8_t* src, uint32_t clen, uint8_t* out, uint32_t isize) {
    // the device reads whole words around the stream: give the copy the same slack, at an odd alignment
    uint8_t* buf = (uint8_t*)calloc(clen + 64, 1);
    const int skew = (int)(clen % 4);
    memcpy(buf + 8 + skew, src, clen);
    zinf::Scratch* S = (zinf::Scratch*)calloc(1, sizeof(zinf::Scratch));
    g_bar = nullptr;
    const int rc = zinf::inflate_member<1>(0, 1u, buf + 8 + skew, clen, out, isize, S);
    free(S);
    free(buf);
    return rc;
}

struct LaneArg {
    int lane, w;
    const uint8_t* src;
    uint32_t clen;
    uint8_t* out;
    uint32_t isize;
    zinf::Scratch* S;
    int rc;
};

template <int W>
static void* lane_main(void* p) {
    LaneArg* a = (LaneArg*)p;
    a->rc = zinf::inflate_member<W>(a->lane, (1u << W) - 1u, a->src, a->clen, a->out, a->isize, a->S);
    return nullptr;
}

// returns the lanes' common result code, or -100 if they disagree.  `out` needs isize + 64 bytes.
extern "C" int zi_inflate_host_lanes(int w, const uint8_t* src, uint32_t clen, uint8_t* out, uint32_t isize) {
    if (w != 2 && w != 4 && w != 8) return -101;
    uint8_t* buf = (uint8_t*)calloc(clen + 64, 1);
    const int skew = (int)((clen + 1) % 4);
    memcpy(buf + 8 + skew, src, clen);
    zinf::Scratch* S = (zinf::Scratch*)calloc(1, sizeof(zinf::Scratch));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, (unsigned)w);
    g_bar = &bar;
    pthread_t th[8];
    LaneArg arg[8];
    for (int l = 0; l < w; l++) {
        arg[l] = LaneArg{l, w, buf + 8 + skew, clen, out, isize, S, 0};
        pthread_create(&th[l], nullptr, w == 2 ? lane_main<2> : (w == 4 ? lane_main<4> : lane_main<8>), &arg[l]);
    }
    for (int l = 0; l < w; l++) pthread_join(th[l], nullptr);
    g_bar = nullptr;
    pthread_barrier_destroy(&bar);
    int rc = arg[0].rc;
    for (int l = 1; l < w; l++)
        if (arg[l].rc != rc) rc = -100;
    free(S);
    free(buf);
    return rc;
}
