// host_tables.h -- tables the kernels and the host resolver share, built once per context.
#pragma once

#include <stdint.h>

#include <vector>

#include "device_types.h"

namespace b200 {

// Mode S CRC-24 (generator 0xFFF409) and the syndrome-repair tables of crc.c.
class CrcTables {
  public:
    explicit CrcTables(int nfix);

    // modesChecksum (crc.c:67-82): remainder over the first bits-24 bits XOR the last 24 bits
    uint32_t checksum(const uint8_t *msg, int bits) const;
    // modesChecksumDiagnose (crc.c:389-412): nullptr = not repairable; errors == 0 for syndrome 0
    const ErrorInfo *diagnose(uint32_t syndrome, int bits) const;
    // modesChecksumFix (crc.c:417-425)
    static void fix(uint8_t *msg, const ErrorInfo *ei);

    const uint32_t *bit_syndromes() const { return bit_syndrome_; } // 112 entries, crc.c:59-64
    const std::vector<ErrorInfo> &short_table() const { return short_; }
    const std::vector<ErrorInfo> &long_table() const { return long_; }

  private:
    void build(int bits, int max_correct, int max_detect, std::vector<ErrorInfo> &out) const;
    uint32_t byte_table_[256];
    bool pos_ready_ = false;
    uint32_t pos_table_[14][256]; // syndrome of byte value b at byte position i of a 112-bit frame
    uint32_t bit_syndrome_[112];
    std::vector<ErrorInfo> short_, long_;
    ErrorInfo no_errors_;
};

// init_uc8_lookup (convert.c:35-61): table[I | Q << 8] for the little-endian u16 a uc8 sample is
void build_uc8_table(uint16_t *table65536);
// init_sc16q11_lookup (convert.c:270-294) of a build with -DSC16Q11_TABLE_BITS=bits: 1 << (2 * bits) entries
void build_sc16q11_table(int bits, uint16_t *table);

} // namespace b200
