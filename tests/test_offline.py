"""SURVEY.md section 8(f) N4 -- the offline dataset darkener (utils/lowlight_process.py:10-87) on the GPU: nvJPEG decode ->
``apply_lowlight`` (the library's truncating uint8 writer) -> nvJPEG encode.  Parity is stated on the PRE-ENCODE uint8 arrays
(bit-exact against the reference's expression on the same decoded pixels); the files themselves are checked for names,
shapes and JPEG-level closeness only (decoders / encoders of different libraries are not byte-identical)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _write_inputs(d):
    from torchvision.io import encode_jpeg, encode_png

    bus = torch.from_numpy(np.load(os.path.join(GOLD, "bus640.npz"))["u8"])   # [3,640,640] uint8 RGB
    assert bus.shape == (3, 640, 640)
    imgs = {"a.jpg": bus, "b.JPG": bus[:, 100:420, 60:540].contiguous(), "c.jpeg": bus[:, ::2, ::2].contiguous(),
            "d.png": bus[:, 300:377, 200:333].contiguous()}
    for name, im in imgs.items():
        data = encode_png(im) if name.endswith(".png") else encode_jpeg(im, quality=97)
        with open(os.path.join(d, name), "wb") as f:
            f.write(data.numpy().tobytes())
    with open(os.path.join(d, "notes.txt"), "w") as f:
        f.write("not an image")
    return imgs


def test_darken_directory_matches_reference_expression(tmp_path):
    from torchvision.io import ImageReadMode, decode_image

    import dedark_yolo_b200 as dd
    from dedark_yolo_b200 import offline

    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    imgs = _write_inputs(str(src))
    dev = torch.device("cuda:0")
    p = 5.0                                                     # utils/lowlight_process.py:94 (__main__)
    n0 = dd.launch_count()
    assert dd.darken_directory(str(src), str(dst), lowlight_param=p, batch_size=3, device=dev) == 4
    assert dd.launch_count() - n0 == 2                          # one darkening launch per batch of files (4 files, batch 3)
    assert sorted(os.listdir(dst)) == sorted(imgs)              # same file names, the .txt is ignored

    # pre-encode parity: the decoded pixels through the library vs the reference's expression evaluated by torch on the same GPU
    names = sorted(imgs)
    decoded = offline.decode_rgb([str(src / n) for n in names], dev)
    dark = offline.darken_images(decoded, p)
    for n, u8, q in zip(names, decoded, dark):
        assert u8.dtype == torch.uint8 and u8.shape == imgs[n].shape
        ref = (torch.pow(u8.float() / 255, p) * 255).to(torch.uint8)    # (img * 255).astype(np.uint8): truncation
        assert torch.equal(q, ref), n
        # the GPU decoder against the host decoder on the same file: JPEG decoders differ by a level or two (IDCT rounding,
        # chroma upsampling: measured mean |diff| 1.1 levels on the bus image)
        host = decode_image(torch.from_numpy(np.fromfile(str(src / n), dtype=np.uint8)), mode=ImageReadMode.RGB)
        assert float((host.float() - u8.cpu().float()).abs().mean()) < 2.5
        # the written file decodes to something close to the pre-encode array (quality 95)
        back = decode_image(torch.from_numpy(np.fromfile(str(dst / n), dtype=np.uint8)), mode=ImageReadMode.RGB)
        assert back.shape == u8.shape
        assert float((back.float() - q.cpu().float()).abs().mean()) < (0.01 if n.endswith(".png") else 3.0)


def test_darken_directory_needs_cuda(tmp_path):
    import dedark_yolo_b200 as dd

    with pytest.raises(RuntimeError):
        dd.darken_directory(str(tmp_path), str(tmp_path / "o"), device="cpu")
