"""CPU oracle for the CosyVoice2 hot path (TEST INFRASTRUCTURE ONLY).

This package is the checker for the CUDA product in ``cosyvoice_b200``.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path never
routes through it.

Contents
--------
* ``refimport``   – import shims that make the *unmodified* reference modules
                    under ``/root/reference`` importable in the build container
                    (used only to pin the restatement and to generate goldens).
* ``weights``     – deterministic synthetic state_dicts keyed by the reference's
                    own state_dict names (no pretrained weights exist offline).
* ``hift``/``flow``/``lm``/``mel``/``sampling`` – plain torch-fp32 restatements
                    of the reference algorithm, each function citing the
                    reference file:line it follows (``lm`` also holds the
                    text-streaming ``inference_bistream`` and the CosyVoice3LM
                    variant).
* ``dit``/``hift_causal``/``model3`` – the CosyVoice3 stack (DiT flow, causal
                    vocoder with the fp64 f0 predictor, CosyVoice3Model glue),
                    pinned the same way; the x_transformers rotary embedding
                    used by the DiT is not installed offline and is restated
                    from its published algorithm (parity unpinned for that
                    piece, see ``refimport._RotaryEmbedding``).

Parity pin: the reference ships no golden vectors.  The
restatement is pinned against outputs of the reference itself, imported in the
build container by ``oracle/make_golden.py``; those outputs are committed under
``tests/golden/`` together with the generating script.  Third-party arithmetic
that is not vendored in the reference tree (transformers Qwen2, diffusers
Attention/GELU, librosa mel filterbank) is restated from its published
algorithm; for those pieces parity is pinned only against the versions
installed in the build container (transformers 5.5.0) or is unpinned
(diffusers 0.29.0, librosa 0.10.2) - stated again in DESIGN.md.
"""
