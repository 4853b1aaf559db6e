"""Device-side image front-end, second half: the reference's centre crop (openvla_utils.py:616-648).
CPU: properties of the numpy restatement (TensorFlow is absent, so no golden image exists: parity of this sub-step is
unpinned and says so).  GPU: the kernel is bit-exact against the restatement, and the engine's uint8 entry point with
the crop switched on equals the same entry point fed CPU-cropped frames."""
import numpy as np
import pytest
import torch

from oracle import image_prep as IP


def test_restatement_properties():
    g = np.random.default_rng(0)
    img = g.integers(0, 256, (224, 224, 3), dtype=np.uint8)
    assert np.array_equal(IP.center_crop_image(img, 1.0), img)                 # the whole image: identity
    for v in (0, 1, 77, 128, 254, 255):                                        # constants survive uint8 -> float -> uint8
        assert np.unique(IP.center_crop_image(np.full((224, 224, 3), v, np.uint8))).tolist() == [v]
    out = IP.center_crop_image(img)
    assert out.shape == (224, 224, 3) and out.dtype == np.uint8
    flipped = IP.center_crop_image(img[::-1, ::-1].copy())[::-1, ::-1]                     # centred: flip-symmetric,
    d = np.abs(flipped.astype(int) - out.astype(int))                                      # up to fp32 rounding of the
    assert d.max() <= 1 and (d == 0).mean() > 0.99                                         # sample positions
    y1, x1, y2, x2 = IP.crop_box(0.9)
    assert abs(float((y2 - y1) * (x2 - x1)) - 0.9) < 1e-6 and abs(float(y1 + y2) - 1.0) < 1e-6   # area 0.9, centred
    # a linear ramp is reproduced by bilinear sampling: output pixel i sits at y1*223 + i*(y2-y1)
    ramp = np.repeat(np.arange(224, dtype=np.float32)[:, None], 224, 1)
    r8 = np.clip(ramp, 0, 255).astype(np.uint8)
    o = IP.center_crop_image(np.stack([r8] * 3, -1))[:, 0, 0].astype(np.float32)
    want = float(y1) * 223 + np.arange(224) * float(y2 - y1)
    assert np.abs(o - want).max() <= 1.0
    small = IP.center_crop_image(g.integers(0, 256, (200, 200, 3), dtype=np.uint8))       # other input sizes resample too
    assert small.shape == (224, 224, 3)


@pytest.mark.parametrize("H,W,scale", [(224, 224, 0.9), (200, 240, 0.9), (256, 320, 0.81)])
def test_restatement_against_an_independent_bilinear_sampler(H, W, scale):
    """TensorFlow is absent, so the restatement cannot be pinned against tf.image.crop_and_resize itself; the next best
    thing is an independent implementation of the same sampling rule: torch's grid_sample with align_corners=True puts
    normalised coordinate c at pixel (c + 1) / 2 * (size - 1), which is crop_and_resize's `y1 * (H - 1) + i * scale`
    grid when the normalised box corners are mapped linearly.  Different code, different fp32 operation order: the two
    must agree to one uint8 step, and almost everywhere exactly."""
    g = np.random.default_rng(5)
    img = g.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ours = IP.center_crop_image(img, scale)
    y1, x1, y2, x2 = (float(v) for v in IP.crop_box(scale))
    n = IP.OPENVLA_IMAGE_SIZE
    t = torch.linspace(0, 1, n, dtype=torch.float64)
    ys = (y1 + (y2 - y1) * t) * 2 - 1          # normalised [0,1] box coordinate -> grid_sample's [-1,1]
    xs = (x1 + (x2 - x1) * t) * 2 - 1
    grid = torch.stack(torch.meshgrid(ys, xs, indexing="ij")[::-1], -1)[None]    # (1, n, n, 2) as (x, y)
    x = torch.from_numpy(img).permute(2, 0, 1)[None].double() / 255.0
    v = torch.nn.functional.grid_sample(x, grid, mode="bilinear", padding_mode="border", align_corners=True)
    ref = (v.clamp(0, 1) * 255.5).floor().clamp(0, 255)[0].permute(1, 2, 0).numpy().astype(np.int64)
    d = np.abs(ref - ours.astype(np.int64))
    assert d.max() <= 1 and (d == 0).mean() > 0.995, (d.max(), (d == 0).mean())


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,scale", [(224, 224, 0.9), (224, 224, 0.5), (200, 200, 0.9), (256, 320, 0.81)])
def test_device_crop_is_bit_exact_against_the_restatement(H, W, scale):
    from vla_adapter_b200 import ops

    g = np.random.default_rng(1)
    imgs = g.integers(0, 256, (5, H, W, 3), dtype=np.uint8)
    imgs[0] = 255
    imgs[1, ::2] = 0
    out = ops.center_crop_u8(torch.from_numpy(imgs).cuda(), scale).cpu().numpy()
    ref = np.stack([IP.center_crop_image(im, scale) for im in imgs])
    assert out.shape == ref.shape
    assert np.array_equal(out, ref), f"{(out != ref).sum()} of {out.size} bytes differ"


@pytest.mark.gpu
def test_engine_crops_on_the_device():
    from oracle import vla_oracle as O
    from vla_adapter_b200.engine import VLAEngine

    cfg = O.OracleConfig(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=512, pro=False)
    W = O.make_weights(cfg, seed=23)
    _, ids, prop = O.make_inputs(cfg, 3, 12, seed=23)
    g = np.random.default_rng(2)
    frames = g.integers(0, 256, (3, 2, 224, 224, 3), dtype=np.uint8)
    cropped = np.stack([[IP.center_crop_image(f) for f in s] for s in frames])
    eng = VLAEngine(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=512, max_batch=3, max_prompt_len=12)
    eng.load_flat(W)
    eng.finalize()
    _, n_plain = eng.predict_action_batch(ids, None, None, prop, images_u8=torch.from_numpy(frames))
    _, n_cpu = eng.predict_action_batch(ids, None, None, prop, images_u8=torch.from_numpy(cropped))
    eng.set_center_crop(0.9)
    _, n_dev = eng.predict_action_batch(ids, None, None, prop, images_u8=torch.from_numpy(frames))
    _, n_dev2 = eng.predict_action_batch(ids, None, None, prop, images_u8=torch.from_numpy(frames))   # graph replay
    eng.set_center_crop(0.0)
    _, n_off = eng.predict_action_batch(ids, None, None, prop, images_u8=torch.from_numpy(frames))
    eng.close()
    assert np.array_equal(n_dev, n_cpu) and np.array_equal(n_dev2, n_cpu)
    assert np.array_equal(n_off, n_plain) and not np.array_equal(n_plain, n_cpu)
