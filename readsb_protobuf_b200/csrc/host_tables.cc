// host_tables.cc -- CRC-24 tables, syndrome-repair tables and the uc8 magnitude table.
#include "host_tables.h"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace b200 {

namespace {
constexpr uint32_t kPoly = 0xfff409u; // crc.c:31
}

CrcTables::CrcTables(int nfix) {
    // crc.c:46-57: remainder of every single byte
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i << 16;
        for (int j = 0; j < 8; ++j)
            c = (c & 0x800000u) ? ((c << 1) ^ kPoly) : (c << 1);
        byte_table_[i] = c & 0xffffffu;
    }
    // crc.c:59-64: syndrome of each single-bit error of a 112-bit frame.  The CRC is linear, so
    // the syndrome of bit i is x^(111-i) mod g for the message part and the bit itself for the
    // parity part; computing it through checksum() keeps one definition.
    uint8_t msg[14];
    memset(msg, 0, sizeof(msg));
    for (int i = 0; i < 112; ++i) {
        msg[i >> 3] = (uint8_t) (0x80u >> (i & 7));
        bit_syndrome_[i] = checksum(msg, 112);
        msg[i >> 3] = 0;
    }
    memset(&no_errors_, 0, sizeof(no_errors_));
    pos_ready_ = false;
    // the CRC is linear: the syndrome of a frame is the XOR of the syndromes of its bytes
    for (int i = 0; i < 14; ++i)
        for (uint32_t b = 0; b < 256; ++b) {
            uint32_t x = 0;
            for (int k = 0; k < 8; ++k)
                if (b & (0x80u >> k))
                    x ^= bit_syndrome_[8 * i + k];
            pos_table_[i][b] = x;
        }
    pos_ready_ = true;

    // crc.c:358-383
    if (nfix == 1) {
        build(56, 1, 1, short_);
        build(112, 1, 1, long_);
    } else if (nfix >= 2) {
        build(56, 2, 4, short_);
        build(112, 2, 4, long_);
    }
}

uint32_t CrcTables::checksum(const uint8_t *msg, int bits) const {
    const int n = bits / 8;
    if (pos_ready_) {
        // same value as the byte-serial form below, without its dependency chain
        const int off = 14 - n; // short frames use the tail of the 112-bit positions (crc.c:143)
        uint32_t x = 0;
        for (int i = 0; i < n; ++i)
            x ^= pos_table_[i + off][msg[i]];
        return x;
    }
    uint32_t rem = 0;
    for (int i = 0; i < n - 3; ++i)
        rem = ((rem << 8) ^ byte_table_[msg[i] ^ ((rem >> 16) & 0xffu)]) & 0xffffffu;
    return rem ^ ((uint32_t) msg[n - 3] << 16) ^ ((uint32_t) msg[n - 2] << 8) ^ (uint32_t) msg[n - 1];
}

// prepareErrorTable (crc.c:184-354).  Patterns are enumerated as explicit nested loops instead of
// the reference's recursion; the result only depends on the set of (syndrome, pattern) pairs.
void CrcTables::build(int bits, int max_correct, int max_detect, std::vector<ErrorInfo> &out) const {
    out.clear();
    const int offset = 112 - bits;
    // bits 0..4 (the DF field) are never repaired, crc.c:214-215
    for (int b0 = 5; b0 < bits; ++b0) {
        ErrorInfo e;
        memset(&e, 0, sizeof(e));
        e.syndrome = bit_syndrome_[b0 + offset];
        e.errors = 1;
        e.bit[0] = (int8_t) b0;
        e.bit[1] = -1;
        out.push_back(e);
        if (max_correct >= 2) {
            for (int b1 = b0 + 1; b1 < bits; ++b1) {
                ErrorInfo e2 = e;
                e2.syndrome ^= bit_syndrome_[b1 + offset];
                e2.errors = 2;
                e2.bit[1] = (int8_t) b1;
                out.push_back(e2);
            }
        }
    }
    std::sort(out.begin(), out.end(), [](const ErrorInfo &x, const ErrorInfo &y) { return x.syndrome < y.syndrome; });

    // crc.c:247-267: a syndrome produced by more than one pattern is not repairable at all
    {
        std::vector<ErrorInfo> uniq;
        uniq.reserve(out.size());
        size_t i = 0;
        while (i < out.size()) {
            size_t j = i + 1;
            while (j < out.size() && out[j].syndrome == out[i].syndrome)
                ++j;
            if (j == i + 1)
                uniq.push_back(out[i]);
            i = j;
        }
        out.swap(uniq);
    }

    // crc.c:269-298: drop entries that a (max_correct+1 .. max_detect)-bit error would alias
    if (max_detect > max_correct) {
        auto find = [&](uint32_t syn) -> ErrorInfo * {
            auto it = std::lower_bound(out.begin(), out.end(), syn,
                                       [](const ErrorInfo &x, uint32_t s) { return x.syndrome < s; });
            return (it != out.end() && it->syndrome == syn) ? &*it : nullptr;
        };
        // every pattern of 1..max_detect bits; only those with more than max_correct bits flag
        std::vector<int> idx; // current combination
        std::vector<uint32_t> syn;
        // iterative depth-first enumeration of combinations in lexicographic order
        idx.push_back(5);
        syn.push_back(0);
        while (!idx.empty()) {
            const int depth = (int) idx.size(); // number of bits in the pattern being formed
            int &i = idx.back();
            if (i >= bits) {
                idx.pop_back();
                syn.pop_back();
                if (!idx.empty())
                    ++idx.back();
                continue;
            }
            const uint32_t s = syn.back() ^ bit_syndrome_[i + offset];
            if (depth > max_correct) {
                ErrorInfo *hit = find(s);
                if (hit)
                    hit->errors = -1;
            }
            if (depth < max_detect) {
                idx.push_back(i + 1);
                syn.push_back(s);
            } else {
                ++i;
            }
        }
        out.erase(std::remove_if(out.begin(), out.end(), [](const ErrorInfo &x) { return x.errors == -1; }), out.end());
    }
}

const ErrorInfo *CrcTables::diagnose(uint32_t syndrome, int bits) const {
    if (syndrome == 0)
        return &no_errors_;
    const std::vector<ErrorInfo> &t = (bits == 56) ? short_ : long_;
    auto it = std::lower_bound(t.begin(), t.end(), syndrome,
                               [](const ErrorInfo &x, uint32_t s) { return x.syndrome < s; });
    return (it != t.end() && it->syndrome == syndrome) ? &*it : nullptr;
}

void CrcTables::fix(uint8_t *msg, const ErrorInfo *ei) {
    for (int i = 0; i < ei->errors; ++i)
        msg[ei->bit[i] >> 3] ^= (uint8_t) (0x80u >> (ei->bit[i] & 7));
}

void build_uc8_table(uint16_t *table) {
    // convert.c:45-58.  The double division and the float arithmetic are written exactly as the
    // reference has them; this file is compiled without FMA contraction.
    for (int i = 0; i <= 255; ++i) {
        for (int q = 0; q <= 255; ++q) {
            float fI = (float) ((i - 127.5) / 127.5);
            float fQ = (float) ((q - 127.5) / 127.5);
            float magsq = fI * fI + fQ * fQ;
            if (magsq > 1)
                magsq = 1;
            float mag = sqrtf(magsq);
            table[i * 256 + q] = (uint16_t) (mag * 65535.0f + 0.5f);
        }
    }
}

void build_sc16q11_table(int bits, uint16_t *table) {
    // convert.c:280-291 with USE_BITS = bits, LOSE_BITS = 11 - bits: a double division rounded to float,
    // then float arithmetic as written there (no FMA contraction in this file)
    const int lose = 11 - bits;
    for (int i = 0; i < 2048; i += (1 << lose)) {
        for (int q = 0; q < 2048; q += (1 << lose)) {
            float fI = (float) (i / 2048.0), fQ = (float) (q / 2048.0);
            float magsq = fI * fI + fQ * fQ;
            if (magsq > 1)
                magsq = 1;
            float mag = sqrtf(magsq);
            table[((unsigned) (i >> lose) << bits) | (unsigned) (q >> lose)] = (uint16_t) (mag * 65535.0f + 0.5f);
        }
    }
}

} // namespace b200
