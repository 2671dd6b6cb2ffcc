"""Golden architecture tables (.tex) written by the UNMODIFIED reference (oracle/_ref/serial: print_architecture_tex,
src/auxil.c:872-1097) through its own Python method print_arch_tex.

Run where /root/reference exists:   python tests/golden/make_golden_arch_tex.py
One file per (network, column selection) of ARCH_TEX_CASES -> tests/golden/arch_tex/<net>_<sel>.tex; the GPU test
tests/test_gpu_network.py::test_print_arch_tex builds the same networks on the B200 core and compares the bytes.
(The reference then shells out to pdflatex, which this image does not have: only the .tex is kept.)
"""
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver as rd  # noqa: E402
from tests import netdefs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# column selections: the method's defaults, everything on, a sparse one
SELECTIONS = {
    "default": {},
    "all": dict(size=1, in_size=1, f_size=1, out_size=1, stride=1, padding=1, in_padding=1, activation=1, bias=1, dropout=1,
                param_count=1),
    "sparse": dict(size=0, in_size=1, f_size=0, out_size=1, stride=0, padding=0, activation=1, param_count=1),
}
NETS = {
    "mini_darknet": lambda: netdefs.mini_darknet(),
    "lenet": lambda: netdefs.lenet(batch=4),
}
ARCH_TEX_CASES = [(n, s) for n in NETS for s in SELECTIONS]


def main():
    out_dir = os.path.join(HERE, "arch_tex")
    os.makedirs(out_dir, exist_ok=True)
    for net_name, make in NETS.items():
        ref = rd.RefNet(make(), "C_BLAS")
        for sel_name, kw in SELECTIONS.items():
            tmp = tempfile.mkdtemp() + "/"
            with rd._Quiet():
                ref.cnn.print_arch_tex(tmp, "arch", network=0, **kw)
            dst = os.path.join(out_dir, "%s_%s.tex" % (net_name, sel_name))
            shutil.copyfile(os.path.join(tmp, "arch.tex"), dst)
            shutil.rmtree(tmp)
            print("wrote", dst)


if __name__ == "__main__":
    main()
