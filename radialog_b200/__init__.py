"""radialog_b200 — B200-native (sm_100a) implementation of RaDialog's image->report inference hot path.

Public surface (mirrors the reference's call sites):
  * ``Blip2Qformer.forward_image(image) -> (q_out, image_embeds)``            (vision.py)
  * ``LlamaForCausalLM`` / ``PeftModelForCausalLM`` with ``generate(...)``     (llm.py)
  * ``Prompter``                                                               (prompter.py)
  * ``ReportPipeline`` — both stages glued with a device-tensor hand-off       (pipeline.py)
Compute lives in ``lib/libradialog_b200.so`` (C-ABI: include/radialog_b200.h), built by ``python -m radialog_b200.build``.
Importing this package does not need a GPU; constructing a model does, and there is no CPU fallback.
"""
from .prompter import Prompter  # noqa: F401
from .synth import LlamaCfg, VisionCfg, tiny_llama_cfg, tiny_vision_cfg  # noqa: F401


def __getattr__(name):
    if name in ("LlamaForCausalLM", "PeftModelForCausalLM", "GreedySearchDecoderOnlyOutput"):
        from . import llm
        return getattr(llm, name)
    if name == "Blip2Qformer":
        from . import vision
        return vision.Blip2Qformer
    if name == "ReportPipeline":
        from . import pipeline
        return pipeline.ReportPipeline
    raise AttributeError(name)
