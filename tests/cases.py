"""Shared parity-case definitions: constructor kwargs, seeds and seeded synthetic inputs.

Used by tests/golden/make_golden.py (reference + oracle, build container only) and by the
tests (oracle vs golden on CPU, CUDA path vs golden/oracle on the GPU).  Weights are never
stored: every module is built under ``torch.manual_seed(seed)`` with torch's default inits, which
consumes the CPU generator identically in the reference and in the product's parameter holders
(checked through the per-tensor digests in the golden files).
"""
import hashlib

import torch

CVIVIT_CASES = {
    # BASELINE.json configs[0]
    "cfg1": dict(
        seed=0,
        ctor=dict(dim=256, codebook_size=65536, image_size=64, patch_size=16, temporal_patch_size=2,
                  spatial_depth=2, temporal_depth=2, use_vgg_and_gan=False),
        video=(1, 3, 5, 64, 64), video_seed=1),
    # rectangular image / patch, odd temporal patch, small heads, 10-bit codebook, batch 2
    "rect": dict(
        seed=3,
        ctor=dict(dim=128, codebook_size=1024, image_size=(32, 48), patch_size=(8, 16),
                  temporal_patch_size=3, spatial_depth=1, temporal_depth=2, dim_head=32, heads=4,
                  use_vgg_and_gan=False),
        video=(2, 3, 7, 32, 48), video_seed=4),
    # single image (4-D input, first-frame path only)
    "image": dict(
        seed=5,
        ctor=dict(dim=64, codebook_size=256, image_size=32, patch_size=8, temporal_patch_size=2,
                  spatial_depth=1, temporal_depth=1, dim_head=32, heads=2, channels=1,
                  use_vgg_and_gan=False),
        video=(3, 1, 32, 32), video_seed=6),
    # lookup_free_quantization=False: the cosine-sim VectorQuantize codebook (cvivit.py:321)
    "cosine_vq": dict(
        seed=7,
        ctor=dict(dim=64, codebook_size=300, image_size=32, patch_size=8, temporal_patch_size=2, spatial_depth=1,
                  temporal_depth=1, dim_head=32, heads=2, use_vgg_and_gan=False, lookup_free_quantization=False),
        video=(2, 3, 5, 32, 32), video_seed=8),
}

MASKGIT_CASES = {
    "small": dict(
        seed=10,
        ctor=dict(dim=128, num_tokens=512, max_seq_len=128, heads=4, dim_head=32, depth=2,
                  dim_context=96),
        batch=2, patch_shape=(3, 4, 4), ctx_len=7, ctx_valid=(7, 4), input_seed=11),
    "wide": dict(  # default head geometry (8 x 64), ragged text lengths
        seed=12,
        ctor=dict(dim=256, num_tokens=1024, max_seq_len=256, depth=1, dim_context=768),
        batch=3, patch_shape=(2, 4, 4), ctx_len=9, ctx_valid=(9, 1, 5), input_seed=13),
}

CRITIC_CASES = {
    "small": dict(
        seed=20,
        ctor=dict(dim=128, num_tokens=512, max_seq_len=128, has_cross_attn=True, heads=4, dim_head=32,
                  depth=2, dim_context=96),
        batch=2, patch_shape=(3, 4, 4), ctx_len=7, ctx_valid=(7, 4), input_seed=21),
}

# Full Phenaki.sample runs (C-ViViT 'rect'-like tokenizer + MaskGit [+ TokenCritic]).
SAMPLE_CASES = {
    "confidence": dict(  # no critic -> 1 - p scores
        seed=30, steps=6, cond_scale=3.0, num_frames=7, batch=2, ctx_len=6, ctx_valid=(6, 3),
        critic=False, prime=False, noise_seed=31),
    "critic_primed": dict(  # token critic + priming with 4 frames, cond_scale 5 (configs[4]-like)
        seed=32, steps=5, cond_scale=5.0, num_frames=6, batch=2, ctx_len=6, ctx_valid=(2, 6),
        critic=True, prime=True, prime_frames=4, noise_seed=33),
}

SAMPLE_CVIVIT = dict(dim=64, codebook_size=256, image_size=(16, 24), patch_size=(8, 8),
                     temporal_patch_size=3, spatial_depth=1, temporal_depth=1, dim_head=32, heads=2,
                     use_vgg_and_gan=False)
SAMPLE_MASKGIT = dict(dim=64, num_tokens=256, max_seq_len=64, heads=2, dim_head=32, depth=2,
                      dim_context=48)
SAMPLE_CRITIC = dict(dim=64, num_tokens=256, max_seq_len=64, has_cross_attn=True, heads=2, dim_head=32,
                     depth=1, dim_context=48)


def seeded_randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def synthetic_text_embeds(batch, ctx_len, dim_context, valid, seed):
    """(b, L, dim_context) N(0,1) rows with rows >= valid[b] zeroed (= T5 padding, t5.py:95-103)."""
    e = seeded_randn((batch, ctx_len, dim_context), seed)
    for b, v in enumerate(valid):
        e[b, v:] = 0.0
    return e


def token_inputs(case, num_tokens):
    g = torch.Generator().manual_seed(case["input_seed"])
    n = 1
    for d in case["patch_shape"]:
        n *= d
    ids = torch.randint(0, num_tokens + 1, (case["batch"], n), generator=g)  # includes mask id
    ctx = synthetic_text_embeds(case["batch"], case["ctx_len"], case["ctor"]["dim_context"],
                                case["ctx_valid"], case["input_seed"] + 1000)
    return ids, ctx


def state_digest(sd):
    """Order-independent digest of a state dict (names, shapes, raw bytes)."""
    h = hashlib.sha256()
    for k in sorted(sd):
        t = sd[k].detach().cpu().contiguous()
        h.update(k.encode())
        h.update(str(tuple(t.shape)).encode())
        h.update(str(t.dtype).encode())
        h.update(t.numpy().tobytes())
    return h.hexdigest()


class NoiseTape:
    """Records / replays the uniform draws of the sampling loop in reference order.

    ``tape(shape, tag)`` draws from a private CPU generator seeded with ``seed`` -- the same
    stream the reference consumed when it ran under ``torch.manual_seed(seed)`` and drew with
    ``tensor.uniform_()`` (phenaki_pytorch.py:70-71, 88-90)."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws = {}

    def __call__(self, shape, tag):
        u = torch.zeros(shape).float().uniform_(0, 1, generator=self.gen)
        self.draws[tag] = u
        return u

# make_video (phenaki_pytorch.py:691-714): three scenes chained through the trailing `prime` frames of the previous
# scene, TokenCritic scoring, default cond_scale 3 -- BASELINE.json configs[4] in miniature (one prompt per scene).
MAKE_VIDEO_CASE = dict(seed=50, steps=4, texts=("a cat", "a dog", "a fish"), num_frames=(7, 6, 6), prime_lengths=4,
                       ctx_len=5, ctx_valid=(5, 3, 4), noise_seed=51)


def make_video_text_table(case):
    """text -> (1, L, dim_context) synthetic T5 embedding (zero rows = padding)."""
    return {t: synthetic_text_embeds(1, case["ctx_len"], SAMPLE_MASKGIT["dim_context"], (v,), case["seed"] + 100 + i)
            for i, (t, v) in enumerate(zip(case["texts"], case["ctx_valid"]))}


# Phenaki.forward training loss + gradients (phenaki_pytorch.py:562-687): the reference under ``torch.manual_seed(noise_seed)``
# draws rand_step, the permutation uniform and (with a critic) the V-wide gumbel uniform, in that order.
TRAIN_CASES = {
    "generator": dict(  # MaskGit only (critic=None): cross entropy at the masked positions
        seed=60, steps=18, maskgit=MASKGIT_CASES["small"]["ctor"], critic=None, batch=2, patch_shape=(3, 4, 4),
        ctx_len=7, ctx_valid=(7, 4), input_seed=61, noise_seed=62),
    "with_critic": dict(  # + TokenCritic: gumbel-sampled predictions -> BCE, loss = ce + critic_loss_weight * bce
        seed=63, steps=6, maskgit=SAMPLE_MASKGIT, critic=SAMPLE_CRITIC, batch=3, patch_shape=(3, 2, 3),
        ctx_len=6, ctx_valid=(6, 1, 4), input_seed=64, noise_seed=65),
    "self_critic": dict(  # self_token_critic=True: Linear(dim,1) on the MaskGit embeddings, both losses reach MaskGit
        seed=66, steps=8, maskgit=SAMPLE_MASKGIT, critic=None, self_critic=True, batch=2, patch_shape=(3, 2, 3),
        ctx_len=5, ctx_valid=(3, 5), input_seed=67, noise_seed=68),
}


def train_inputs(case):
    """ids (b, t, h, w) int64 in [0, V) and synthetic text embeddings."""
    g = torch.Generator().manual_seed(case["input_seed"])
    ids = torch.randint(0, case["maskgit"]["num_tokens"], (case["batch"], *case["patch_shape"]), generator=g)
    ctx = synthetic_text_embeds(case["batch"], case["ctx_len"], case["maskgit"]["dim_context"], case["ctx_valid"],
                                case["input_seed"] + 1000)
    return ids, ctx


# ---- round 2: the paths VERDICT r01 found untested -------------------------------------------------------------------
# Phenaki(self_token_critic=True).sample (phenaki_pytorch.py:307-336, 512-545): SelfCritic scores drive the re-masking.
SELF_CRITIC_SAMPLE_CASE = dict(seed=34, steps=5, cond_scale=3.0, num_frames=7, batch=2, ctx_len=6, ctx_valid=(4, 6),
                               noise_seed=35)

# video_mask (attention.py:164-167 key mask of the self-attention; phenaki_pytorch.py:181-190, 265-302): the first
# `valid[b]` tokens of sequence b are real, the rest padding.  MaskGit 'small' / TokenCritic 'small' (48 tokens).
VIDEO_MASK_VALID = dict(maskgit_small=(48, 29), critic_small=(31, 48))


def video_mask_of(valid, n):
    return torch.arange(n)[None, :] < torch.tensor(valid)[:, None]


# Phenaki.forward(videos, video_frame_mask=...) (phenaki_pytorch.py:587-612, cvivit.py:365-373): raw videos are
# tokenised live, frames -> token mask, masked-subset sampling among the valid tokens, key-masked attention, CE + BCE.
FRAME_MASK_TRAIN_CASE = dict(seed=70, steps=6, maskgit=SAMPLE_MASKGIT, critic=SAMPLE_CRITIC, batch=2,
                             video=(2, 3, 7, 16, 24), frames_valid=(7, 4), patch_shape=(3, 2, 3), ctx_len=6,
                             ctx_valid=(6, 2), input_seed=71, noise_seed=72)


def frame_mask_of(valid, frames):
    return torch.arange(frames)[None, :] < torch.tensor(valid)[:, None]
