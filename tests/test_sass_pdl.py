"""Static check of the built library: programmatic dependent launch is only safe when NO global-memory access of a kernel is
scheduled in front of its griddepcontrol.wait (SASS: ACQBULK).  The source puts pdl_begin() first in every kernel, but ptxas is
free to move non-coherent loads (ld.global.nc / __ldg) across the wait -- it did, once (dd_predictor.cu, fc_fwd_kernel) -- so the
rule is checked on the machine code.  A few kernels read PARAMETERS or data that is final at least two kernels earlier ahead of the
wait on purpose: there every access in front of the wait must be one of those loads (the producer's data is read behind it).
"""
import re
import shutil
import subprocess

import pytest

from dedark_yolo_b200 import _lib

ALLOWED_EARLY = {   # kernel-name fragment -> SASS opcodes allowed in front of ACQBULK
    "fc_fwd_kernel": {"LDG.E.128", "LDG.E.128.CONSTANT", "LDG.E.CONSTANT"},   # weights and biases; the activations use ld.global.cg
    # the batch (x, IcA, A: inputs of the step, older than fc_fwd_kernel, which waits before it releases this kernel): the first
    # segment's per-row columns and first tile; the predictor's output (feat) is read with ld.global.cg behind the wait
    "recovery_fwd_kernel": {"LDG.E.CONSTANT", "LDG.E.128.CONSTANT", "LDG.E.64.CONSTANT", "UTMALDG.3D"},
    # the same in the backward (x, IcA, A and the cotangent g).  Its wait sits inside the segment loop, so in the linear listing the
    # loop-carried flush of the per-plane-strip sums (STG.E, executed from the second segment on) is laid out in front of it.
    # the finalize kernel behind it (the main kernel waits before it releases it): feat (ld.global.cg, first in program order) and
    # the rows' first three columns of x / IcA; the main kernel's partial sums and dx are ld.global.cg behind the wait (volatile
    # asm statements keep their order)
    "recovery_bwd_finalize_kernel": {"LDG.E.CONSTANT", "LDG.E.U16.CONSTANT", "LDG.E.STRONG.GPU"},
    "recovery_bwd_kernel": {"LDG.E.CONSTANT", "LDG.E.128.CONSTANT", "LDG.E.64.CONSTANT", "LDG.E.U16.CONSTANT", "UTMALDG.3D",
                            "LDGSTS.E.BYPASS.128.ZFILL", "LDGDEPBAR", "STG.E"},
    "predictor_tail_kernel": {"LDG.E.CONSTANT"},
    # the prepared weights of the tensor-core convolutions (written by conv1_fwd_prep_kernel, at least two kernels earlier in
    # the stream; every tensor-core kernel and fc_bwd release their dependents only after their own wait) are copied early
    # the resized batch r (an input of the forward pass) of the first tile; the cotangent tile follows behind the wait
    "conv1_wgrad_kernel": {"UTMALDG.3D"},
    "conv_tc_fwdILi32E": {"LDGSTS.E.BYPASS.128", "LDGDEPBAR"},
    "conv_tc_bwd": {"LDGSTS.E.BYPASS.128", "LDGDEPBAR"},
}
GLOBAL_OP = re.compile(r"\*/\s+(?:@!?U?P\d+\s+)?((?:LDG|LD\.E|LDGSTS|UTMALDG|UBLKCP|ATOMG|ATOM\.|RED\.|REDG|STG|ST\.E|UTMASTG)\S*)")


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_no_global_access_in_front_of_griddepcontrol_wait():
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            kernels[fn] = {"wait": False, "early": []}
            continue
        if fn is None:
            continue
        k = kernels[fn]
        if "ACQBULK" in line:
            k["wait"] = True
        elif not k["wait"]:
            g = GLOBAL_OP.search(line)
            if g:
                k["early"].append(g.group(1).rstrip(";"))
    assert len(kernels) > 50
    bad = {}
    for name, k in kernels.items():
        assert k["wait"], f"{name}: no griddepcontrol.wait (every kernel of the library starts with pdl_begin)"
        allowed = next((ops for frag, ops in ALLOWED_EARLY.items() if frag in name), set())
        extra = [op for op in k["early"] if op not in allowed]
        if extra:
            bad[name] = sorted(set(extra))
    assert not bad, f"global-memory access scheduled in front of griddepcontrol.wait: {bad}"
    # the two exceptions read the producer's data with ld.global.cg (STRONG.GPU in SASS), which is never hoisted
    for frag in ALLOWED_EARLY:
        names = [n for n in kernels if frag in n]
        assert names, frag
