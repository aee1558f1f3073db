#!/bin/bash
# Development aid (no GPU needed): the host resolver under AddressSanitizer + UBSan and under ThreadSanitizer --
# the recorded kernel outputs of tests/golden/ (or the dumps given as arguments) through tools/resolver_bench.cc with
# 1, 3 and 8 threads, and the filter-rule property test of tests/cpp/.
#   scripts/sanitize_host.sh [span dumps in stream order ...]
set -e
cd "$(dirname "$0")/.."
dumps=("$@")
[ ${#dumps[@]} -eq 0 ] && dumps=($(ls tests/golden/resolver_span_*.bin | sort -t_ -k3 -n))
src="readsb_protobuf_b200/csrc/resolver.cc readsb_protobuf_b200/csrc/host_tables.cc"
inc="-Iinclude -Ireadsb_protobuf_b200/csrc"
out=$(mktemp -d)
for san in address,undefined thread; do
    g++ -O1 -g -fsanitize=$san -fno-omit-frame-pointer -std=c++17 $inc tools/resolver_bench.cc $src -lpthread -o $out/bench
    g++ -O1 -g -fsanitize=$san -std=c++17 $inc tests/cpp/test_icao_filter.cc $src -lpthread -o $out/rule
    for th in 1 3 8; do
        for mode in optimistic always-optimistic prescan; do
            echo "== -fsanitize=$san, $th threads, $mode"
            B200_RESOLVER_THREADS=$th B200_RESOLVER_PREDICT=$mode B200_RESOLVER_MIN_LIVE=0 B200_RESOLVER_MIN_BLOCKS=1 B200_RESOLVER_MIN_LIVE_PER_RUN=1 RB_REPS=2 \
                $out/bench "${dumps[@]}" 2>&1 | grep -v "^rep" | tail -2
        done
    done
    $out/rule 800 | tail -1
done
rm -rf $out
