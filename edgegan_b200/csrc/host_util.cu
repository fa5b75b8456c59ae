// Host-side helpers of the checkpoint path (edgegan_b200/checkpoint.py): CRC-32C (Castagnoli) as used by the
// TensorFlow tensor-bundle format that tf.train.Saver writes (edgegan/models/edgegan.py:421,547,635-657 save / restore
// through it).  Slicing-by-8, ~1.5 GB/s per core: a 200 MB checkpoint is checksummed in ~0.15 s.
#include "common.cuh"

namespace {

uint32_t g_tab[8][256];
bool g_tab_ready = false;

void init_tab() {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;   // reflected Castagnoli polynomial
        g_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 255u];
    g_tab_ready = true;
}

}  // namespace

extern "C" unsigned int eg_crc32c(const void* data, long long n, unsigned int crc) {
    if (!g_tab_ready) init_tab();
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint32_t c = ~crc;
    while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 7u)) { c = g_tab[0][(c ^ *p++) & 255u] ^ (c >> 8); --n; }
    while (n >= 8) {
        uint64_t w;
        __builtin_memcpy(&w, p, 8);
        w ^= c;
        c = g_tab[7][w & 255u] ^ g_tab[6][(w >> 8) & 255u] ^ g_tab[5][(w >> 16) & 255u] ^ g_tab[4][(w >> 24) & 255u] ^
            g_tab[3][(w >> 32) & 255u] ^ g_tab[2][(w >> 40) & 255u] ^ g_tab[1][(w >> 48) & 255u] ^ g_tab[0][(w >> 56) & 255u];
        p += 8; n -= 8;
    }
    while (n > 0) { c = g_tab[0][(c ^ *p++) & 255u] ^ (c >> 8); --n; }
    return ~c;
}
