/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product, never linked into it.
 *
 * ref_netfmt: runs the UNMODIFIED reference's Beast and raw output writers
 * (modesSendBeastOutput net_io.c:769-835, modesSendRawOutput net_io.c:870-896) over the messages of a
 * result file written by ref_demod.  The writers are static functions, so this translation unit
 * includes the reference's own net_io.c where it lies under /root/reference (-I$(REF), nothing is
 * copied) and calls them with an in-memory net_writer; what the file needs from the protobuf-c
 * runtime and from anet.c is never reached and is stubbed below.
 *
 * usage: ref_netfmt --in RESULT --beast-out FILE --raw-out FILE [--mlat] [--no-verbatim]
 */
#include "net_io.c"

#include <limits.h>

const char protobuf_c_empty_string[] = "";
struct ProtobufCMessageDescriptor { int unused; };

struct _Modes Modes;

/* ---- never reached by the two writers ---- */
#define STUB_PACK(prefix, type)                                                                 \
    size_t prefix##__get_packed_size(const type *m) { (void) m; return 0; }                     \
    size_t prefix##__pack(const type *m, uint8_t *out) { (void) m; (void) out; return 0; }
STUB_PACK(aircrafts_update, AircraftsUpdate)
STUB_PACK(receiver, Receiver)
STUB_PACK(statistics, Statistics)
const ProtobufCMessageDescriptor aircrafts_update__descriptor, statistic_entry__descriptor;
void statistics__polar_range_entry__init(Statistics__PolarRangeEntry *m) { (void) m; }
void aircraft_history__init(AircraftHistory *m) { (void) m; }
const ProtobufCMessageDescriptor aircraft_meta__descriptor, aircraft_meta__nav_modes__descriptor,
    aircraft_meta__valid_source__descriptor, receiver__descriptor;
void aircraft_meta__init(AircraftMeta *m) { static const AircraftMeta v = AIRCRAFT_META__INIT; *m = v; }
void aircraft_meta__nav_modes__init(AircraftMeta__NavModes *m) { static const AircraftMeta__NavModes v = AIRCRAFT_META__NAV_MODES__INIT; *m = v; }
void aircraft_meta__valid_source__init(AircraftMeta__ValidSource *m) { static const AircraftMeta__ValidSource v = AIRCRAFT_META__VALID_SOURCE__INIT; *m = v; }
void receiver__init(Receiver *m) { static const Receiver v = RECEIVER__INIT; *m = v; }

/* ---- result file records (oracle/ref_harness.c) ---- */
#pragma pack(push, 1)
struct result_header { char magic[4]; uint32_t version; uint64_t n_msgs, n_blocks, n_samples; };
struct result_msg {
    uint64_t timestampMsg, sysTimestampMsg;
    double signalLevel;
    uint32_t crc, addr;
    int32_t score;
    uint8_t msgbits, msgtype, correctedbits, reserved;
    uint8_t msg[14], verbatim[14];
};
#pragma pack(pop)
#define RESULT_STATS_BYTES 136

int main(int argc, char **argv) {
    const char *in = NULL, *beast_path = NULL, *raw_path = NULL;
    int mlat = 0, verbatim = 1;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--in") && i + 1 < argc) in = argv[++i];
        else if (!strcmp(argv[i], "--beast-out") && i + 1 < argc) beast_path = argv[++i];
        else if (!strcmp(argv[i], "--raw-out") && i + 1 < argc) raw_path = argv[++i];
        else if (!strcmp(argv[i], "--mlat")) mlat = 1;
        else if (!strcmp(argv[i], "--no-verbatim")) verbatim = 0;
        else { fprintf(stderr, "ref_netfmt: bad argument %s\n", argv[i]); return 2; }
    }
    if (!in || !beast_path || !raw_path) {
        fprintf(stderr, "usage: ref_netfmt --in RESULT --beast-out FILE --raw-out FILE [--mlat] [--no-verbatim]\n");
        return 2;
    }
    FILE *f = fopen(in, "rb");
    struct result_header hdr;
    if (!f || fread(&hdr, sizeof (hdr), 1, f) != 1 || memcmp(hdr.magic, "MDSR", 4) || fseek(f, RESULT_STATS_BYTES, SEEK_CUR)) {
        fprintf(stderr, "ref_netfmt: cannot read %s\n", in);
        return 2;
    }
    struct result_msg *msgs = calloc(hdr.n_msgs ? hdr.n_msgs : 1, sizeof (*msgs));
    if (fread(msgs, sizeof (*msgs), hdr.n_msgs, f) != hdr.n_msgs) {
        fprintf(stderr, "ref_netfmt: short result file\n");
        return 2;
    }
    fclose(f);

    memset(&Modes, 0, sizeof (Modes));
    Modes.net_verbatim = (int8_t) verbatim;
    Modes.mlat = (int8_t) mlat;
    Modes.net_output_flush_size = INT_MAX; /* completeWrite never flushes: we drain the buffer ourselves */

    /* writers with one pretend connection, so that prepareWrite hands out the buffer (net_io.c:733-749) */
    struct net_service beast_service, raw_service;
    struct net_writer beast_writer;
    memset(&beast_service, 0, sizeof (beast_service));
    memset(&raw_service, 0, sizeof (raw_service));
    memset(&beast_writer, 0, sizeof (beast_writer));
    beast_service.connections = raw_service.connections = 1;
    beast_service.writer = &beast_writer;
    raw_service.writer = &Modes.raw_out;
    beast_writer.data = malloc(MODES_OUT_BUF_SIZE);
    beast_writer.service = &beast_service;
    Modes.raw_out.data = malloc(MODES_OUT_BUF_SIZE);
    Modes.raw_out.service = &raw_service;

    FILE *fb = fopen(beast_path, "wb"), *fr = fopen(raw_path, "wb");
    if (!fb || !fr) {
        perror("ref_netfmt");
        return 2;
    }
    for (uint64_t i = 0; i < hdr.n_msgs; ++i) {
        struct modesMessage mm;
        memset(&mm, 0, sizeof (mm));
        mm.msgbits = msgs[i].msgbits;
        mm.msgtype = msgs[i].msgtype;
        mm.timestampMsg = msgs[i].timestampMsg;
        mm.signalLevel = msgs[i].signalLevel;
        memcpy(mm.msg, msgs[i].msg, 14);
        memcpy(mm.verbatim, msgs[i].verbatim, 14);
        modesSendBeastOutput(&mm, &beast_writer);
        fwrite(beast_writer.data, 1, (size_t) beast_writer.dataUsed, fb);
        beast_writer.dataUsed = 0;
        modesSendRawOutput(&mm);
        fwrite(Modes.raw_out.data, 1, (size_t) Modes.raw_out.dataUsed, fr);
        Modes.raw_out.dataUsed = 0;
    }
    fclose(fb);
    fclose(fr);
    return 0;
}
