// netfmt.cc -- the wire formats readsb's Beast and raw output services use for a message
// (SURVEY.md 8f row 1): host-side, batched over the message list a process call produced.
//
//   b200_format_beast  replaces modesSendBeastOutput (net_io.c:769-835)
//   b200_format_raw    replaces modesSendRawOutput   (net_io.c:870-896)
//
// The formatting is a few byte moves per message; it lives on the host next to the message list (the
// records are already there for the resolver) rather than in a kernel.
#include <math.h>
#include <stdio.h>

#include "readsb_b200.h"

namespace {

// bounded output cursor: counts every byte, stores those that fit
struct Out {
    uint8_t *p;
    uint64_t cap, at;
    void put(uint8_t b) {
        if (at < cap)
            p[at] = b;
        ++at;
    }
    void put_escaped(uint8_t b) { // Beast framing: a literal 0x1a is sent twice
        put(b);
        if (b == 0x1a)
            put(b);
    }
};

} // namespace

extern "C" uint64_t b200_format_beast(const b200_message *msgs, uint64_t n, int net_verbatim, uint8_t *out, uint64_t cap) {
    Out o{out, out ? cap : 0, 0};
    for (uint64_t i = 0; i < n; ++i) {
        const b200_message &mm = msgs[i];
        const int len = mm.msgbits / 8;
        const uint8_t *msg = net_verbatim ? mm.verbatim : mm.msg; // Modes.net_verbatim, net_io.c:775
        uint8_t type;
        switch (len) { // net_io.c:781-789: anything else is not sent
            case 7: type = '2'; break;
            case 14: type = '3'; break;
            case 2: type = '1'; break;
            default: continue;
        }
        o.put(0x1a);
        o.put(type);
        for (int shift = 40; shift >= 0; shift -= 8) // 12 MHz timestamp, 48 bits, big-endian
            o.put_escaped((uint8_t) (mm.timestampMsg >> shift));
        int sig = (int) round(sqrt(mm.signalLevel) * 255); // net_io.c:817-821
        if (mm.signalLevel > 0 && sig < 1)
            sig = 1;
        if (sig > 255)
            sig = 255;
        o.put_escaped((uint8_t) sig);
        for (int j = 0; j < len; ++j)
            o.put_escaped(msg[j]);
    }
    return o.at;
}

extern "C" uint64_t b200_format_raw(const b200_message *msgs, uint64_t n, int net_verbatim, int mlat, char *out, uint64_t cap) {
    static const char hex[] = "0123456789ABCDEF";
    Out o{reinterpret_cast<uint8_t *>(out), out ? cap : 0, 0};
    for (uint64_t i = 0; i < n; ++i) {
        const b200_message &mm = msgs[i];
        const int len = mm.msgbits / 8;
        const uint8_t *msg = net_verbatim ? mm.verbatim : mm.msg;
        if (mlat && mm.timestampMsg) { // net_io.c:879-884: "@" + 12 hex digits of the 12 MHz timestamp
            char head[16];
            const int nh = snprintf(head, sizeof(head), "@%012llX", (unsigned long long) mm.timestampMsg);
            for (int j = 0; j < nh; ++j)
                o.put((uint8_t) head[j]);
        } else {
            o.put('*');
        }
        for (int j = 0; j < len; ++j) {
            o.put((uint8_t) hex[msg[j] >> 4]);
            o.put((uint8_t) hex[msg[j] & 15]);
        }
        o.put(';');
        o.put('\n');
    }
    return o.at;
}
