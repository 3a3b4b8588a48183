"""Helper processes of tests/test_checkpoint_roundtrip.py (run as a script, never collected by pytest).

  write <path>   vanilla reference + ``integrate.install()``: build the drop-in (seed 0), attach parse_model's attributes,
                 pickle it the way trainer.py:408-433 does (whole object)
  read <path> <out.pt>
                 a process in which ``dedark_yolo_b200`` CANNOT be imported: un-pickle, run the module on the CPU
"""
import os
import pickle
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, path = sys.argv[1], sys.argv[2]
    extra = sys.argv[3:]
    sys.argv = sys.argv[:1]
    from baseline import reference_runtime as R
    if mode == "write":
        R.load_ultralytics()
        import dedark_yolo_b200.integrate as it
        it.install()
        from ultralytics.nn.modules.llie import lowlight_recovery
        torch.manual_seed(0)
        m = lowlight_recovery(3)
        m.i, m.f, m.type, m.np = 0, -1, "ultralytics.nn.modules.llie.lowlight_recovery", 164943
        with open(path, "wb") as f:
            pickle.dump({"model": m.half()}, f)
        print("written by", type(m).__module__, type(m.extractor).__module__)
    else:
        sys.modules["dedark_yolo_b200"] = None  # any import of the drop-in package raises ImportError
        R.load_ultralytics()
        with open(path, "rb") as f:
            m = pickle.load(f)["model"].float().eval()
        assert type(m).__module__ == "ultralytics.nn.modules.llie" and type(m.extractor).__module__ == "ultralytics.nn.modules.common"
        assert type(m.filters[0]).__module__ == "ultralytics.nn.modules.filtersB" and (m.i, m.f, m.np) == (0, -1, 164943)
        g = torch.Generator().manual_seed(3)
        x = torch.rand(2, 3, 40, 48, generator=g)
        with torch.no_grad():
            y = m(x)
        torch.save({"x": x, "y": y, "state": m.state_dict()}, extra[0])
        print("read as", type(m).__module__)


if __name__ == "__main__":
    main()
