#!/bin/bash
# CPU sanitizer pass over everything of this repo that runs on the host without CUDA (compute-sanitizer is closed on the
# GPU pool): a scratch copy of the tree under build/san/ with
#   * the golden model, its decoder and the reference-driver simulation (oracle/, oracle/refsim/ + kernel/cedar.c and
#     userspace/h264enc.c from /root/reference) built with AddressSanitizer + UndefinedBehaviorSanitizer,
#   * the host build of the product's entropy / EPB logic (tests/host_harness.cpp over csrc/entropy.cuh, h264_core.cuh),
#   * the CLI's parallel batch reader (csrc/h264enc.c, self-test build) under ThreadSanitizer and ASan,
# and the CPU test suite run on top of them.  Usage: tools/sanitize_cpu.sh [report file]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-$ROOT/profiles/r02_sanitizers_cpu.txt}
S=$ROOT/build/san
rm -rf "$S"; mkdir -p "$S"
(cd "$ROOT" && tar --exclude=.git --exclude=gpurun_out --exclude=build --exclude='*.o' --exclude=_ref --exclude=__pycache__ \
    --exclude=.pytest_cache --exclude=oracle/liboracle_h264.so --exclude=tests/libhost_harness.so -cf - .) | (cd "$S" && tar xf -)
SAN="-fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer"
cd "$S"
make -s -C oracle CC="gcc $SAN" CFLAGS="-O1 -g -Wall -Wextra -fPIC -std=gnu11" all
g++ $SAN -O1 -g -Wall -Wno-unknown-pragmas -fPIC -shared -std=c++17 -o tests/libhost_harness.so tests/host_harness.cpp
touch tests/libhost_harness.so oracle/liboracle_h264.so
{
echo "# CPU sanitizer pass ($(date -u +%F), $(gcc --version | head -1))"
echo "# flags: $SAN; scratch copy of the tree, the product's CUDA library untouched (it is not involved)"
echo "## pytest -m 'not gpu' on the sanitized golden model / decoder / reference-driver simulation / host harness"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1
PRE="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
# (tests/test_bench_contract.py is left out: it times a 1080p CPU encode, which a sanitized build cannot finish in its limit)
LD_PRELOAD="$PRE" python -m pytest tests/test_oracle.py tests/test_refsim.py tests/test_headers.py tests/test_entropy_host.py \
    tests/test_abi.py tests/test_cli.py tests/test_synth_partition.py -x -q -s -m "not gpu" -p no:cacheprovider 2>&1 | grep -a "runtime error\|AddressSanitizer\|^    #[0-4] \| passed\| failed\|^FAILED\|Error" | head -40
echo "## the oracle's encoder program and the reference CLI on the simulated driver, sanitized, 7 frames 96x80"
LD_PRELOAD="$PRE" python - <<'PY'
import sys; sys.path.insert(0, "tests")
from common import make_clip
open("build_clip.nv12", "wb").write(make_clip("synth", 96, 80, 7).tobytes())
PY
./oracle/_ref/h264enc_sim build_clip.nv12 96 80 build_ref.264 > /dev/null && echo "h264enc_sim rc=0, $(stat -c %s build_ref.264) bytes"
unset LD_PRELOAD
echo "## parallel batch reader of the CLI (h264enc.c, self-test build): ThreadSanitizer, then ASan+UBSan"
head -c 3000017 /dev/urandom > build_rs.bin
for san in "-fsanitize=thread" "$SAN"; do
    gcc $san -O1 -g -std=gnu11 -DH264ENC_READER_SELFTEST -Iinclude -pthread -o build_reader cedarx_h264_encoder_b200/csrc/h264enc.c
    for t in 1 3 8 16; do ./build_reader build_rs.bin 4099 64 $t 2>/dev/null | cmp -n $((3000017 / 4099 * 4099)) - build_rs.bin; done
    echo "reader self-test clean under: $san"
done
} > "$OUT" 2>&1 || { cat "$OUT"; exit 1; }
cat "$OUT"
