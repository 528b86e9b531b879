#!/bin/bash
# AddressSanitizer + UBSan over the HOST side of the library: every test that runs against the CPU test double
# (tests/mock_build.py) is run with the library built -fsanitize=address,undefined and the ASan runtime preloaded
# into python.  No GPU needed.  Output -> profiles/r02_asan_host_suite.txt
cd "$(dirname "$0")/.."
export MXD_MOCK_SANITIZE=address,undefined
python -c "from tests import mock_build; print(mock_build.build('address,undefined'))"
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
ASAN_OPTIONS="detect_leaks=0:halt_on_error=0:abort_on_error=0" UBSAN_OPTIONS="print_stacktrace=1" \
  timeout 3000 python -m pytest tests/test_digest_service.py tests/test_client_flows.py tests/test_cli.py -m "not gpu" -q -x 2>&1 | tee /tmp/asan_suite.$$ | tail -5
echo "AddressSanitizer reports: $(grep -c 'ERROR: AddressSanitizer' /tmp/asan_suite.$$)"
echo "UBSan reports: $(grep -c 'runtime error:' /tmp/asan_suite.$$)"
grep -A12 'ERROR: AddressSanitizer\|runtime error:' /tmp/asan_suite.$$ | head -80
rm -f /tmp/asan_suite.$$
