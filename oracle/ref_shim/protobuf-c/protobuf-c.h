/*
 * TEST INFRASTRUCTURE ONLY (oracle scaffolding) -- not part of the product.
 *
 * Minimal stand-in for <protobuf-c/protobuf-c.h>, which is not installed in this
 * image.  The reference's readsb.h:86 includes it and readsb.h:229 includes the
 * generated readsb.pb-c.h, but the IQ->magnitude->demod->CRC path only needs the
 * *types* those headers declare.  With this header on the include path the
 * reference's own .c files compile unmodified from /root/reference (see
 * oracle/Makefile, target `ref`).
 */
#ifndef ORACLE_SHIM_PROTOBUF_C_H
#define ORACLE_SHIM_PROTOBUF_C_H

#include <stddef.h>
#include <stdint.h>

#define PROTOBUF_C_VERSION_NUMBER 1003003
#define PROTOBUF_C_MIN_COMPILER_VERSION 1000000
#define PROTOBUF_C__BEGIN_DECLS
#define PROTOBUF_C__END_DECLS
#define PROTOBUF_C__FORCE_ENUM_TO_BE_INT_SIZE(tag) , _##tag##_FORCE_INT_SIZE = 0x7fffffff
#define PROTOBUF_C_MESSAGE_INIT(descriptor) { descriptor, 0, NULL }

typedef int protobuf_c_boolean;

typedef struct ProtobufCMessageDescriptor ProtobufCMessageDescriptor;
typedef struct ProtobufCEnumDescriptor ProtobufCEnumDescriptor;
typedef struct ProtobufCAllocator ProtobufCAllocator;
typedef struct ProtobufCBuffer ProtobufCBuffer;

typedef struct ProtobufCMessage {
    const ProtobufCMessageDescriptor *descriptor;
    unsigned n_unknown_fields;
    void *unknown_fields;
} ProtobufCMessage;

typedef struct ProtobufCBinaryData {
    size_t len;
    uint8_t *data;
} ProtobufCBinaryData;

extern const char protobuf_c_empty_string[];

#endif
