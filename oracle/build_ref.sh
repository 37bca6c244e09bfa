#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE. Compiles the reference's own CPU path, from its
# sources where they lie (never copied into this repo), plus oracle/ref_driver.cpp, into
# oracle/_ref/libjpegref.so. The reference's IDE build files are not used.
# Usage: oracle/build_ref.sh [/root/reference]
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${1:-/root/reference}"
src="$ref/src"
if [ ! -f "$src/decoder.cpp" ]; then
    echo "build_ref: reference sources not found under $src (expected on the GPU box: the prebuilt .so travels)" >&2
    exit 3
fi
mkdir -p "$here/_ref"
# -DUSE_CPU_ONLY selects the CPU IDCT/colour branch (decoder.cpp:11 has the define commented out);
# -DNDEBUG turns the reference's int3 asserts off (macro.h:24-31). -O2 as in the reference's Release target.
g++ -std=c++11 -O2 -fPIC -shared -DNDEBUG -DUSE_CPU_ONLY -w \
    -I"$src" \
    "$here/ref_driver.cpp" \
    "$src/bitstream.cpp" "$src/huffman.cpp" "$src/cpuIDCT8x8.cpp" "$src/decoder.cpp" "$src/parser.cpp" \
    -o "$here/_ref/libjpegref.so"
echo "built $here/_ref/libjpegref.so"

# Drop-in proof: the reference's OWN main.cpp + parser.cpp (+ the bit reader / trie units its self
# tests use), unmodified, linked against this repo's decoder shim instead of decoder.cpp,
# cpuIDCT8x8.cpp and oclDCT8x8.cpp. Needs ocljpegdecoder_b200/lib/libb2j.so (built by `make`).
root="$(cd "$here/.." && pwd)"
if [ -f "$root/ocljpegdecoder_b200/lib/libb2j.so" ]; then
    g++ -std=c++11 -O2 -DNDEBUG -w -DB2J_USE_REFERENCE_HEADERS -I"$src" \
        "$src/main.cpp" "$src/parser.cpp" "$src/bitstream.cpp" "$src/huffman.cpp" \
        "$root/ocljpegdecoder_b200/csrc/refshim/decoder_b2j.cpp" \
        -L"$root/ocljpegdecoder_b200/lib" -lb2j -Wl,-rpath,'$ORIGIN/../../ocljpegdecoder_b200/lib' \
        -o "$here/_ref/ocljpegdec_b2j"
    echo "built $here/_ref/ocljpegdec_b2j (reference main+parser on the B200 decoder shim)"
    # The secondary boundary: the reference's own decoder.cpp too, unmodified and WITHOUT USE_CPU_ONLY (its
    # device branch: CPU Huffman, then the ten clidct_* calls of idct.h:9-18), linked against this repo's
    # clidct shim instead of oclDCT8x8.cpp.
    g++ -std=c++11 -O2 -DNDEBUG -w -DB2J_USE_REFERENCE_HEADERS -I"$src" \
        "$src/main.cpp" "$src/parser.cpp" "$src/bitstream.cpp" "$src/huffman.cpp" "$src/decoder.cpp" "$src/cpuIDCT8x8.cpp" \
        "$root/ocljpegdecoder_b200/csrc/refshim/idct_b2j.cpp" \
        -L"$root/ocljpegdecoder_b200/lib" -lb2j -Wl,-rpath,'$ORIGIN/../../ocljpegdecoder_b200/lib' \
        -o "$here/_ref/ocljpegdec_b2jidct"
    echo "built $here/_ref/ocljpegdec_b2jidct (reference main+parser+decoder on the B200 clidct shim)"
fi
