// Host-side helpers of the C ABI that are not GPU work.
//   taco_crc32c: CRC-32C (Castagnoli), slicing-by-8 — the checksum of TensorFlow's tensor-bundle checkpoints
//   (core/lib/hash/crc32c.h); used by tf_checkpoint.py when reference checkpoints are imported / exported.
#include "../../include/taco_capi.h"
#include <cstdint>
#include <cstring>

namespace {
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? 0x82F63B78u : 0u);
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; i++)
            for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
    }
};
const CrcTables g_crc;
}  // namespace

extern "C" uint32_t taco_crc32c(const void* data, size_t n, uint32_t crc) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint32_t c = crc ^ 0xFFFFFFFFu;
    while (n >= 8) {
        uint64_t w;
        std::memcpy(&w, p, 8);
        w ^= c;                                     // little-endian hosts only (x86-64 / aarch64-le)
        c = g_crc.t[7][w & 0xFF] ^ g_crc.t[6][(w >> 8) & 0xFF] ^ g_crc.t[5][(w >> 16) & 0xFF] ^ g_crc.t[4][(w >> 24) & 0xFF] ^
            g_crc.t[3][(w >> 32) & 0xFF] ^ g_crc.t[2][(w >> 40) & 0xFF] ^ g_crc.t[1][(w >> 48) & 0xFF] ^ g_crc.t[0][(w >> 56) & 0xFF];
        p += 8; n -= 8;
    }
    while (n--) c = g_crc.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
