"""adapter/BatchedFHEHIPPIE_b200.cpp: the reference's class over the GPU library, compiled against the reference's
unmodified BatchedFHEHIPPIE.hpp and a minimal lbcrypto shim (OpenFHE itself is absent from the image).
CPU: it builds (where /root/reference exists) and refuses to run without a device.  GPU: the reference's own
known-answer scenario through it ("Matches" exactly twice), on one device and on a device list."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "adapter")
REF_HEADER = "/root/reference/src/Common/Crypto/PrivateIndexedEqualityCheck/BatchedFHEHIPPIE.hpp"


def _binary():
    if os.path.exists(REF_HEADER):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    path = os.path.join(ADAPTER, "test_adapter")
    if not os.path.exists(path):
        pytest.skip("adapter/test_adapter is not built (the reference header is only present in the build container)")
    return path


def test_adapter_builds_against_reference_header_and_has_no_cpu_path():
    if not os.path.exists(REF_HEADER):
        pytest.skip("reference tree not present")
    path = _binary()
    src = open(os.path.join(ADAPTER, "BatchedFHEHIPPIE_b200.cpp")).read()
    assert '#include "BatchedFHEHIPPIE.hpp"' in src and "class BatchedFHEHIPPIE" not in src   # the reference's header, not a copy
    if os.path.exists("/dev/nvidia0"):
        return
    r = subprocess.run([path], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and ("no CPU path" in r.stderr or "CUDA" in r.stderr), r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0", "0,0,0", "all"])
@pytest.mark.parametrize("tables", ["openfhe", "derived"])
def test_adapter_known_answer(devices, tables):
    path = _binary()
    env = dict(os.environ)
    if devices != "all":
        env["PSI_B200_DEVICES"] = devices
    else:
        env.pop("PSI_B200_DEVICES", None)
    r = subprocess.run([path] + (["derived"] if tables == "derived" else []), capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Matches\n") == 2


# ---- the non-batched operator: adapter/FHEHIPPIE_b200.cpp against the reference's unmodified FHEHIPPIE.hpp --------------
REF_HEADER_NB = "/root/reference/src/Common/Crypto/PrivateIndexedEqualityCheck/FHEHIPPIE.hpp"


def _binary_nb():
    if os.path.exists(REF_HEADER_NB):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    path = os.path.join(ADAPTER, "test_adapter_nb")
    if not os.path.exists(path):
        pytest.skip("adapter/test_adapter_nb is not built (the reference header is only present in the build container)")
    return path


def test_nonbatched_adapter_builds_against_reference_header_and_has_no_cpu_path():
    if not os.path.exists(REF_HEADER_NB):
        pytest.skip("reference tree not present")
    path = _binary_nb()
    src = open(os.path.join(ADAPTER, "FHEHIPPIE_b200.cpp")).read()
    assert '#include "FHEHIPPIE.hpp"' in src and "class FHEHIPPIE" not in src   # the reference's header, not a copy
    if os.path.exists("/dev/nvidia0"):
        return
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 2 and ("no CPU path" in r.stderr or "CUDA" in r.stderr), r.stderr


@pytest.mark.gpu
def test_nonbatched_adapter_known_answer():
    """Two PIEs on one context through the reference's class: the one holding the client's element matches exactly
    once, the other never; result limbs identical to the oracle's FHEHIPPIE::run."""
    r = subprocess.run([_binary_nb()], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Matches (PIE 0)\n") == 1 and "Matches (PIE 1)" not in r.stdout
    assert "limb parity with the oracle: identical" in r.stdout
