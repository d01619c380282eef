"""Builds and runs the C++ unit tests of the core (csrc/tests/test_core.cpp), optionally
under AddressSanitizer + UBSan (the reference has no sanitizer targets)."""
import os
import subprocess


def test_native_core_unit_tests():
    from tools import build_native

    exe = build_native.build_cpp_tests()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


def test_native_core_under_asan_ubsan():
    from tools import build_native

    exe = build_native.build_cpp_tests(sanitize=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr[-4000:]
    assert "0 failed" in r.stdout


def test_server_and_client_loopback_under_asan_ubsan():
    """Server + client over loop-back TCP in one sanitized native binary: store round trips,
    eviction, dead writers, garbage on the wire, checkpoint / resume."""
    from tools import build_native

    exe = build_native.build_loopback_test(sanitize=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0:protect_shadow_gap=0",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-6000:]
    assert "0 failed" in r.stdout
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr


def test_server_and_client_loopback_under_tsan():
    """The same binary under ThreadSanitizer: the reactor thread, the async completion thread
    and the client threads of the loop-back scenarios race-free."""
    from tools import build_native

    exe = build_native.build_loopback_test(sanitize="thread")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0:second_deadlock_stack=1")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-6000:]
    assert "0 failed" in r.stdout
    assert "WARNING: ThreadSanitizer" not in r.stderr
