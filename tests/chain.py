"""Oracle replay of make_video (phenaki_pytorch.py:691-714): scenes chained through priming.  Test infrastructure
(used by tests/golden/make_golden.py to pin the oracle against the reference, and by the CPU test suite)."""
import torch

from oracle import phenaki_oracle as O
from tests import cases as C


def oracle_make_video(case, cv_sd, mg_sd, cr_sd, table, tape):
    image_size, patch_size = C.SAMPLE_CVIVIT["image_size"], C.SAMPLE_CVIVIT["patch_size"]
    pt = C.SAMPLE_CVIVIT["temporal_patch_size"]
    hh, ww = image_size[0] // patch_size[0], image_size[1] // patch_size[1]
    n_scenes = len(case["texts"])
    primes = (case["prime_lengths"],) * (n_scenes - 1) + (0,)
    scenes, prime = [], None
    with torch.no_grad():
        for text, nf, next_prime in zip(case["texts"], case["num_frames"], primes):
            prime_ids, pf = None, 0
            if prime is not None:
                prime_ids = O.cvivit_codebook_ids(prime, cv_sd, image_size, patch_size).reshape(1, -1)
                pf = prime.shape[2]
            num_tokens = (hh * ww) * ((nf - 1) // pt + 1) if prime is None else (hh * ww) * (nf // pt)
            patch_shape = (1 + (nf + pf - 1) // pt, hh, ww)
            ids = O.sample_token_ids(mg_sd, num_tokens=num_tokens, patch_shape=patch_shape, batch=1,
                                     steps=case["steps"], heads=C.SAMPLE_MASKGIT["heads"], text_embeds=table[text],
                                     prime_ids=prime_ids, cond_scale=3.0, critic_sd=cr_sd, noise_fn=tape)
            full = ids if prime_ids is None else torch.cat((prime_ids, ids), dim=-1)
            v = O.cvivit_decode_from_ids(full, cv_sd, image_size, patch_size)[:, :, pf:]
            scenes.append(v)
            prime = v[:, :, -next_prime:]
    return torch.cat(scenes, dim=2), scenes
