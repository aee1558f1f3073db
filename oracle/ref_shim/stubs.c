/*
 * TEST INFRASTRUCTURE ONLY (oracle scaffolding) -- not part of the product.
 *
 * Link-time stand-ins for the parts of readsb that cannot be built in this image
 * (net_io.c and readsb.pb-c.c need the protobuf-c runtime, interactive.c needs
 * ncurses).  None of them is on the IQ->magnitude->demod->CRC path.  Compiled
 * against the reference's own headers from /root/reference; see oracle/Makefile.
 *
 * modesQueueOutput() is the one stub with behaviour: with Modes.net=1 and
 * Modes.net_verbatim=1, useModesMessage() (mode_s.c:2146-2173) forwards *every*
 * accepted message to it, so it is where the harness captures the reference's
 * decoded modesMessage records.
 */
#include "readsb.h"

const char protobuf_c_empty_string[] = "";

struct ProtobufCMessageDescriptor { int unused; };
const ProtobufCMessageDescriptor aircraft_meta__descriptor;
const ProtobufCMessageDescriptor aircraft_meta__nav_modes__descriptor;
const ProtobufCMessageDescriptor aircraft_meta__valid_source__descriptor;
const ProtobufCMessageDescriptor receiver__descriptor;

void aircraft_meta__init(AircraftMeta *m) {
    static const AircraftMeta init_value = AIRCRAFT_META__INIT;
    *m = init_value;
}

void aircraft_meta__nav_modes__init(AircraftMeta__NavModes *m) {
    static const AircraftMeta__NavModes init_value = AIRCRAFT_META__NAV_MODES__INIT;
    *m = init_value;
}

void aircraft_meta__valid_source__init(AircraftMeta__ValidSource *m) {
    static const AircraftMeta__ValidSource init_value = AIRCRAFT_META__VALID_SOURCE__INIT;
    *m = init_value;
}

void receiver__init(Receiver *m) {
    static const Receiver init_value = RECEIVER__INIT;
    *m = init_value;
}

void cleanupNetwork(void) {}
void generateAircraftProtoBuf(void) {}
void generateHistoryProtoBuf(const char *f) { (void) f; }
void generateReceiverProtoBuf(void) {}
void generateStatsProtoBuf(void) {}
void interactiveCleanup(void) {}
void interactiveInit(void) {}
void interactiveShowData(void) {}
void modesInitNet(void) {}
void modesNetPeriodicWork(void) {}
void modesNetSecondWork(void) {}

/* capture hook, set by the harness (NULL in the stock readsb_ref binary) */
void (*oracle_capture_hook)(struct modesMessage *mm) = NULL;

void modesQueueOutput(struct modesMessage *mm, struct aircraft *a) {
    (void) a;
    if (oracle_capture_hook)
        oracle_capture_hook(mm);
}
