"""Host-side mirror of the reference's predict_action interface on top of libvla_b200.so.

`VLAEngine.predict_action` keeps the signature and return tuple of
OpenVLAForActionPrediction.predict_action (prismatic/extern/hf/modeling_prismatic.py:892-972);
`predict_action_batch` is the batched form the reference lacks (it hard-codes bs=1 at MP:855, 871).
Weights come from the reference's own three state dicts (HF model, action head, proprio projector).
All arithmetic runs in the CUDA library; this file only prepares indices and moves pointers.  There is no
CPU fallback: without the library or a CUDA device construction fails."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _lib, tokens

NUM_TOKENS = tokens.NUM_TOKENS
LLM_DIM = 896
_DTYPES = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}

# prismatic/vla/constants.py:28-54
PLATFORM_CONSTANTS = {
    "LIBERO": dict(chunk_len=8, action_dim=7, proprio_dim=8, normalization="bounds_q99"),
    "CALVIN": dict(chunk_len=8, action_dim=7, proprio_dim=8, normalization="bounds_q99"),
    "ALOHA": dict(chunk_len=25, action_dim=14, proprio_dim=14, normalization="bounds"),
    "BRIDGE": dict(chunk_len=5, action_dim=7, proprio_dim=7, normalization="bounds_q99"),
}


class VLAEngine:
    def __init__(self, n_images: int = 2, chunk_len: int = 8, action_dim: int = 7, proprio_dim: int = 8,
                 pro: bool = False, dino_depth: int = 24, siglip_depth: int = 27, llm_layers: int = 24,
                 vocab_size: int = 151936, max_batch: int = 1, max_prompt_len: int = 64, causal: bool = True,
                 norm_stats: Optional[Dict[str, Dict[str, Any]]] = None, normalization: str = "bounds_q99",
                 device: int | str = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("VLAEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self._out_bufs = {}
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        torch.cuda.set_device(self.device)
        self.n_images, self.chunk_len, self.action_dim, self.proprio_dim = n_images, chunk_len, action_dim, proprio_dim
        self.pro, self.max_batch, self.max_prompt_len = pro, max_batch, max_prompt_len
        self.num_patches = 256 * n_images
        self.norm_stats = norm_stats
        if normalization not in ("bounds", "bounds_q99"):
            raise ValueError("Unsupported action/proprio normalization type detected!")  # MP:797
        self.normalization = normalization
        self._stats_key = None
        cfg = _lib.VlaCfg(n_images, chunk_len, action_dim, proprio_dim, int(pro), dino_depth, siglip_depth,
                          llm_layers, vocab_size, max_batch, max_prompt_len, int(causal))
        h = C.c_void_p()
        rc = self.lib.vla_create(C.byref(cfg), C.byref(h))
        self._h = h
        if rc != 0:
            try:
                _lib.check(rc, self._h if self._h else None)
            finally:
                if self._h:
                    self.lib.vla_destroy(self._h)
                    self._h = None
        self._finalized = False

    # ------------------------------------------------------------------ weights
    def load_tensor(self, name: str, t: torch.Tensor) -> None:
        if t.dtype not in _DTYPES:
            raise ValueError(f"unsupported dtype {t.dtype} for {name}")
        t = t.detach().contiguous()
        shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
        _lib.check(self.lib.vla_load_tensor(self._h, name.encode(), t.data_ptr(), _DTYPES[t.dtype], t.dim(), shape),
                   self._h)

    def load_state_dicts(self, vla: Dict[str, torch.Tensor], action_head: Dict[str, torch.Tensor],
                         proprio_projector: Dict[str, torch.Tensor]) -> None:
        """Takes the reference's three state dicts (openvla_utils.py:272-327, 412-453, 482-539); a DDP
        `module.` prefix is stripped like openvla_utils.py:230-250 does."""
        for prefix, sd in (("vla.", vla), ("head.", action_head), ("proprio.", proprio_projector)):
            for k, v in sd.items():
                if k.startswith("module."):
                    k = k[len("module."):]
                if not torch.is_tensor(v) or not v.is_floating_point():
                    continue
                self.load_tensor(prefix + k, v)

    def load_flat(self, weights: Dict[str, torch.Tensor]) -> None:
        """Weights already keyed with the vla./head./proprio. prefixes."""
        for k, v in weights.items():
            self.load_tensor(k, v)

    @classmethod
    def from_reference_modules(cls, vla, action_head, proprio_projector, **kw) -> "VLAEngine":
        """Builds an engine from live reference modules (OpenVLAForActionPrediction, L1RegressionActionHead,
        ProprioProjector), reading every shape parameter from them."""
        head_model = action_head.model
        pro = hasattr(head_model.mlp_resnet_blocks[0], "k_self")
        vb = vla.vision_backbone
        tc = vla.config.text_config
        eng = cls(n_images=vb.get_num_images_in_input(), action_dim=action_head.action_dim,
                  chunk_len=kw.pop("chunk_len", 8), proprio_dim=proprio_projector.proprio_dim, pro=pro,
                  dino_depth=len(vb.featurizer.blocks), siglip_depth=len(vb.fused_featurizer.blocks),
                  llm_layers=tc.num_hidden_layers, vocab_size=vla.get_input_embeddings().weight.shape[0],
                  norm_stats=getattr(vla, "norm_stats", None), **kw)
        eng.load_state_dicts(vla.state_dict(), action_head.state_dict(), proprio_projector.state_dict())
        eng.finalize()
        return eng

    # ------------------------------------------------------------------ statistics (MP:977-1001)
    @staticmethod
    def _check_unnorm_key(norm_stats, unnorm_key):
        if unnorm_key is None:
            assert len(norm_stats) == 1, (
                f"Your model was trained on more than one dataset, please pass a `unnorm_key` from the following "
                f"options to choose the statistics used for un-normalizing actions: {norm_stats.keys()}")
            unnorm_key = next(iter(norm_stats.keys()))
        assert unnorm_key in norm_stats, (
            f"The `unnorm_key` you chose is not in the set of available dataset statistics, "
            f"please choose from: {norm_stats.keys()}")
        return unnorm_key

    def get_action_stats(self, unnorm_key=None):
        unnorm_key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        return self.norm_stats[unnorm_key]["action"]

    def _bounds(self, unnorm_key):
        st = self.get_action_stats(unnorm_key)
        if self.normalization == "bounds":
            mask = st.get("mask", np.ones_like(st["min"], dtype=bool))
            hi, lo = np.array(st["max"]), np.array(st["min"])
        else:
            mask = st.get("mask", np.ones_like(st["q01"], dtype=bool))
            hi, lo = np.array(st["q99"]), np.array(st["q01"])
        return np.asarray(mask, dtype=bool), hi.astype(np.float64), lo.astype(np.float64)

    def _unnormalize_actions(self, normalized_actions: np.ndarray, unnorm_key=None) -> np.ndarray:
        """Exactly MP:799-803, in float64 numpy on the host (56 numbers); the device also writes an fp32
        copy for C-ABI callers."""
        mask, hi, lo = self._bounds(unnorm_key)
        return np.where(mask, 0.5 * (normalized_actions + 1) * (hi - lo + 1e-8) + lo, normalized_actions)

    def _push_stats(self, unnorm_key):
        if self.norm_stats is None:
            return
        key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        if key == self._stats_key:
            return
        mask, hi, lo = self._bounds(key)
        A = self.action_dim
        if not (len(hi) == len(lo) == len(mask) == A):
            raise ValueError(f"action statistics of `{key}` have {len(hi)} dims, engine was built for {A}")
        _lib.check(self.lib.vla_set_action_stats(self._h, (C.c_double * A)(*hi), (C.c_double * A)(*lo),
                                                 (C.c_uint8 * A)(*[int(m) for m in mask])), self._h)
        self._stats_key = key

    def finalize(self) -> None:
        if self.norm_stats is not None and len(self.norm_stats) == 1:
            self._push_stats(None)
        _lib.check(self.lib.vla_finalize(self._h), self._h)
        self._finalized = True

    # ------------------------------------------------------------------ forward
    def _prep(self, input_ids, attention_mask):
        """-> (ext_ids (B, L+65) int64, aq_index (B, L+65) int32, prompt_len (B) int32 or None).

        Prompts of different lengths come RIGHT-padded with their attention mask (rows of ones followed by zeros, the
        tokenizer's padding_side="right"), or as a list of 1-D id tensors.  Every sample gets the reference's own bs=1
        preamble (tokens.build on its un-padded ids, MP:922-937) and the result is padded on the right, where causal
        attention cannot see it; `prompt_len` tells the engine where each sample's ActionQuery rows sit."""
        if isinstance(input_ids, (list, tuple)) and len(input_ids) and torch.as_tensor(input_ids[0]).dim() == 1:
            rows = [torch.as_tensor(r).to("cpu", torch.int64) for r in input_ids]
        else:
            ids = torch.as_tensor(input_ids).to("cpu", torch.int64)
            if ids.dim() != 2:
                raise ValueError("input_ids must be (B, L)")
            if attention_mask is None or bool(torch.as_tensor(attention_mask).bool().all()):
                ext, labels, mask, aq_index, _ = tokens.build(ids, None, self.action_dim)
                return ext.contiguous(), aq_index.contiguous(), None
            am = torch.as_tensor(attention_mask).to("cpu").bool()
            if am.shape != ids.shape:
                raise ValueError("attention_mask must have the shape of input_ids")
            lens = am.sum(1)
            if bool((lens < 1).any()) or not bool((am == (torch.arange(ids.shape[1])[None] < lens[:, None])).all()):
                raise ValueError("padded prompts must be right-padded (mask = ones then zeros) and non-empty")
            rows = [ids[b, : int(lens[b])] for b in range(ids.shape[0])]
        L = max(int(r.numel()) for r in rows)
        if min(int(r.numel()) for r in rows) < 1:
            raise ValueError("empty prompt")
        ext = torch.zeros((len(rows), L + NUM_TOKENS + 1), dtype=torch.int64)   # pad id 0: any valid id will do
        aq = torch.full((len(rows), L + NUM_TOKENS + 1), -1, dtype=torch.int32)
        for b, r in enumerate(rows):
            e1, _, _, a1, _ = tokens.build(r[None], None, self.action_dim)
            ext[b, : e1.shape[1]] = e1[0]
            aq[b, : a1.shape[1]] = a1[0]
        lens = torch.tensor([int(r.numel()) for r in rows], dtype=torch.int32)
        if bool((lens == L).all()):
            return ext, aq, None
        return ext, aq, lens

    def predict_device(self, pixel_values: torch.Tensor, ext_ids: torch.Tensor, aq_index: torch.Tensor,
                       proprio: torch.Tensor, want_last_ha: bool = False, prompt_len: Optional[torch.Tensor] = None):
        """Inputs already on the device (bf16 pixels, int64 ids, int32 indices, fp32 proprio[, int32 prompt lengths]);
        enqueues on the current stream and returns device tensors (normalized, unnormalized[, last_ha])."""
        B, Lext = ext_ids.shape
        L = Lext - NUM_TOKENS - 1
        T, A = self.chunk_len, self.action_dim
        if pixel_values.dtype != torch.bfloat16 or not pixel_values.is_contiguous() or \
                tuple(pixel_values.shape) != (B, 6 * self.n_images, 224, 224):
            raise ValueError(f"pixel_values must be contiguous bf16 ({B}, {6 * self.n_images}, 224, 224), got "
                             f"{pixel_values.dtype} {tuple(pixel_values.shape)}")
        for name, t, dt, shape in (("ext_ids", ext_ids, torch.int64, (B, Lext)), ("aq_index", aq_index, torch.int32, (B, Lext)),
                                   ("proprio", proprio, torch.float32, (B, self.proprio_dim))):
            if t.dtype != dt or tuple(t.shape) != shape or not t.is_contiguous() or not t.is_cuda:
                raise ValueError(f"{name} must be a contiguous CUDA {dt} tensor of shape {shape}, got {t.dtype} "
                                 f"{tuple(t.shape)} on {t.device}")
        if not pixel_values.is_cuda:
            raise ValueError("predict_device takes CUDA tensors (use predict_host for host buffers)")
        if prompt_len is not None and (prompt_len.dtype != torch.int32 or tuple(prompt_len.shape) != (B,) or
                                       not prompt_len.is_cuda or not prompt_len.is_contiguous()):
            raise ValueError(f"prompt_len must be a contiguous CUDA int32 tensor of shape ({B},)")
        # The engine replays a CUDA graph keyed on (B, L, buffer addresses): write into buffers that stay put and
        # hand fresh copies (a few hundred bytes per sample) to the caller.
        key = (B, want_last_ha)
        bufs = self._out_bufs.get(key)
        if bufs is None:
            bufs = (torch.empty((B, T, A), dtype=torch.float32, device=self.device),
                    torch.empty((B, T, A), dtype=torch.float32, device=self.device),
                    torch.empty((B, NUM_TOKENS, LLM_DIM), dtype=torch.bfloat16, device=self.device)
                    if want_last_ha else None)
            self._out_bufs[key] = bufs
        out_n, out_u, ha = bufs
        rc = self.lib.vla_predict(self._h, pixel_values.data_ptr(), ext_ids.data_ptr(), aq_index.data_ptr(),
                                  prompt_len.data_ptr() if prompt_len is not None else None,
                                  proprio.data_ptr(), B, L, out_n.data_ptr(), out_u.data_ptr(),
                                  ha.data_ptr() if ha is not None else None,
                                  torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, self._h)
        return out_n.clone(), out_u.clone(), (ha.clone() if ha is not None else None)

    def predict_host(self, pixel_values: torch.Tensor, ext_ids: torch.Tensor, aq_index: torch.Tensor,
                     proprio: torch.Tensor, out_norm: torch.Tensor, out_unnorm: torch.Tensor,
                     out_last_ha: Optional[torch.Tensor] = None, prompt_len: Optional[torch.Tensor] = None) -> None:
        """End-to-end call on HOST tensors (ideally pinned): H2D, forward, D2H, stream sync inside."""
        B, Lext = ext_ids.shape
        if pixel_values.dtype != torch.bfloat16 or tuple(pixel_values.shape) != (B, 6 * self.n_images, 224, 224):
            raise ValueError(f"pixel_values must be bf16 ({B}, {6 * self.n_images}, 224, 224), got {pixel_values.dtype} "
                             f"{tuple(pixel_values.shape)}")
        self._check_host_io(B, Lext, ext_ids, aq_index, proprio, out_norm, out_unnorm, out_last_ha, prompt_len)
        pixel_values = pixel_values.contiguous()
        rc = self.lib.vla_predict_host(self._h, pixel_values.data_ptr(), ext_ids.data_ptr(), aq_index.data_ptr(),
                                       prompt_len.data_ptr() if prompt_len is not None else None,
                                       proprio.data_ptr(), B, Lext - NUM_TOKENS - 1, out_norm.data_ptr(),
                                       out_unnorm.data_ptr(),
                                       out_last_ha.data_ptr() if out_last_ha is not None else None,
                                       torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, self._h)

    def predict_host_u8(self, images_u8: torch.Tensor, ext_ids: torch.Tensor, aq_index: torch.Tensor,
                        proprio: torch.Tensor, out_norm: torch.Tensor, out_unnorm: torch.Tensor,
                        out_last_ha: Optional[torch.Tensor] = None, prompt_len: Optional[torch.Tensor] = None) -> None:
        """predict_host from uint8 frames (B, n_images, 224, 224, 3), HWC, already resized / centre-cropped: the
        processor's ToTensor + Normalize + bf16 cast happen on the device (bit-identical to the CPU path)."""
        B, Lext = ext_ids.shape
        if images_u8.dtype != torch.uint8 or tuple(images_u8.shape) != (B, self.n_images, 224, 224, 3):
            raise ValueError(f"images must be uint8 ({B}, {self.n_images}, 224, 224, 3), got {images_u8.dtype} "
                             f"{tuple(images_u8.shape)}")
        self._check_host_io(B, Lext, ext_ids, aq_index, proprio, out_norm, out_unnorm, out_last_ha, prompt_len)
        images_u8 = images_u8.contiguous()
        rc = self.lib.vla_predict_host_u8(self._h, images_u8.data_ptr(), ext_ids.data_ptr(), aq_index.data_ptr(),
                                          prompt_len.data_ptr() if prompt_len is not None else None,
                                          proprio.data_ptr(), B, Lext - NUM_TOKENS - 1, out_norm.data_ptr(),
                                          out_unnorm.data_ptr(),
                                          out_last_ha.data_ptr() if out_last_ha is not None else None,
                                          torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, self._h)

    def _check_host_io(self, B, Lext, ext_ids, aq_index, proprio, out_norm, out_unnorm, out_last_ha,
                       prompt_len=None) -> None:
        """The C ABI takes raw pointers: every buffer it will read or write is checked here for dtype, size and
        contiguity (the reference raises on mismatched batches, MP:519-522)."""
        T, A = self.chunk_len, self.action_dim

        def need(name, t, dtype, shape):
            if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous() or t.is_cuda:
                raise ValueError(f"{name} must be a contiguous host {dtype} tensor of shape {tuple(shape)}, got "
                                 f"{t.dtype} {tuple(t.shape)}")

        need("ext_ids", ext_ids, torch.int64, (B, Lext))
        need("aq_index", aq_index, torch.int32, (B, Lext))
        need("proprio", proprio, torch.float32, (B, self.proprio_dim))
        need("out_norm", out_norm, torch.float32, (B, T, A))
        need("out_unnorm", out_unnorm, torch.float32, (B, T, A))
        if out_last_ha is not None:
            need("out_last_ha", out_last_ha, torch.bfloat16, (B, NUM_TOKENS, LLM_DIM))
        if prompt_len is not None:
            need("prompt_len", prompt_len, torch.int32, (B,))

    def set_image_norm(self, mean, std) -> None:
        """mean / std of the two backbones, each (2, 3): row 0 DINOv2, row 1 SigLIP (preprocessor_config.json)."""
        import ctypes as C
        m = (C.c_float * 6)(*[float(v) for row in mean for v in row])
        sd = (C.c_float * 6)(*[float(v) for row in std for v in row])
        _lib.check(self.lib.vla_set_image_norm(self._h, m, sd), self._h)

    def set_center_crop(self, crop_scale: float = 0.9) -> None:
        """uint8 frames are centre-cropped on the device (the reference's center_crop_image, OU:616-648) before the
        patch gather; 0 switches it off.  Applies to predict_host_u8 / predict_action_batch(images_u8=...)."""
        _lib.check(self.lib.vla_set_center_crop(self._h, float(crop_scale)), self._h)

    def predict_action_batch(self, input_ids, attention_mask=None, pixel_values=None, proprio=None,
                             unnorm_key=None, return_hidden: bool = False, images_u8=None):
        """(B, L) ids + (B, 6n, 224, 224) pixels + (B, P) proprio -> un-normalised actions (B, T, A) float64
        and normalised actions (B, T, A) float32 [+ last-layer ActionQuery states (B, 1, 64, D) bf16]."""
        if not self._finalized:
            raise RuntimeError("engine not finalized")
        ext, aq, lens = self._prep(input_ids, attention_mask)
        B = ext.shape[0]
        if (pixel_values is None) == (images_u8 is None):
            raise ValueError("pass exactly one of pixel_values (normalised, like the reference) or images_u8")
        pix = torch.as_tensor(pixel_values if images_u8 is None else images_u8)
        if pix.shape[0] != B:
            raise ValueError("Non-homogenous batch of (text, image) input -- forward() does not support mixed batches!")
        if images_u8 is None:
            pix = pix.to(torch.bfloat16).contiguous()
        pr = torch.as_tensor(np.asarray(proprio.cpu() if torch.is_tensor(proprio) else proprio, dtype=np.float32))
        pr = pr.reshape(B, -1).contiguous()
        if pr.shape[1] != self.proprio_dim:
            raise ValueError(f"proprio must have {self.proprio_dim} dims per sample")
        if self.norm_stats is not None:
            self._push_stats(unnorm_key)
        elif unnorm_key is not None:
            # the reference would fail in _check_unnorm_key (MP:980-990): there is nothing to look the key up in
            raise ValueError(f"unnorm_key={unnorm_key!r} given but the engine has no normalisation statistics")
        T, A = self.chunk_len, self.action_dim
        out_n = torch.empty((B, T, A), dtype=torch.float32)
        out_u = torch.empty((B, T, A), dtype=torch.float32)
        ha = torch.empty((B, NUM_TOKENS, LLM_DIM), dtype=torch.bfloat16) if return_hidden else None
        if images_u8 is None:
            self.predict_host(pix.cpu(), ext, aq, pr, out_n, out_u, ha, lens)
        else:
            self.predict_host_u8(pix.cpu(), ext, aq, pr, out_n, out_u, ha, lens)
        normalized = out_n.numpy()
        if self.norm_stats is not None:
            actions = self._unnormalize_actions(normalized.astype(np.float32), unnorm_key)
        else:  # no statistics: the engine's table is the identity (hi = 1, lo = -1): NORMALISED actions come back
            actions = out_u.numpy().astype(np.float64)
        if return_hidden:
            return actions, normalized, ha.view(B, 1, NUM_TOKENS, LLM_DIM)
        return actions, normalized

    def predict_action(self, input_ids=None, unnorm_key=None, proprio=None, proprio_projector=None, action_head=None,
                       noisy_action_projector=None, use_film: bool = False, **kwargs):
        """Drop-in for OpenVLAForActionPrediction.predict_action (MP:892-972), bs=1: returns
        (actions (T, A) float64 ndarray, actions_hidden_states (1, 1, 64, D) bf16 tensor on the device).
        `action_head` / `proprio_projector` are accepted for signature compatibility; their weights were
        taken at construction.  FiLM and the discrete-token branch are not part of the accelerated path."""
        if use_film:
            raise NotImplementedError("use_film=True is outside the accelerated path (default False everywhere)")
        if proprio is None:
            raise ValueError("the L1-regression path needs proprio (AH:53)")
        pixel_values = kwargs["pixel_values"]
        attention_mask = kwargs.get("attention_mask")
        ids = torch.as_tensor(input_ids)
        if ids.dim() != 2 or ids.shape[0] != 1:
            raise ValueError("predict_action is the reference's bs=1 entry point; use predict_action_batch")
        actions, _, ha = self.predict_action_batch(ids, attention_mask, pixel_values, np.asarray(
            proprio.detach().cpu().float() if torch.is_tensor(proprio) else proprio, dtype=np.float32).reshape(1, -1),
            unnorm_key, return_hidden=True)
        return actions[0], ha.to(self.device)

    # ------------------------------------------------------------------ debug / accounting
    def tap(self, name: str) -> torch.Tensor:
        n = C.c_size_t(0)
        _lib.check(self.lib.vla_get_tap(self._h, name.encode(), None, 0, C.byref(n)), self._h)
        buf = torch.empty(n.value // 2, dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.vla_get_tap(self._h, name.encode(), buf.data_ptr(), n.value, C.byref(n)), self._h)
        return buf

    def last_launch_count(self) -> int:
        return int(self.lib.vla_last_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.vla_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
