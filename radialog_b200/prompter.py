"""Prompt-template helper with the call surface of the reference's ``utils/prompter.py:10-50``.

Behaviour kept: ``Prompter(template_name)`` resolves ``data/templates/<name>.json`` relative to the CWD, an empty name
means "alpaca", a missing template raises ``ValueError("Can't read ...")``; ``generate_prompt(instruction, input, label)``
fills ``prompt_input`` / ``prompt_no_input`` and appends the label; ``get_response(output)`` returns the text after the
LAST ``response_split`` marker (multi-turn prompts), stripped.  Added: the one template this path uses (vicuna_v11) is
packaged, so the class also works outside a checkout of the reference.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

_DEFAULT_NAME = "alpaca"
_TEMPLATE_DIR = os.path.join("data", "templates")
_PACKAGED: Dict[str, Dict[str, str]] = {
    "vicuna_v11": dict(description="Template used for Vicuna v1.1. preparation already in code.",
                       prompt_input="{instruction} {input}", prompt_no_input="{instruction}", response_split="ASSISTANT:"),
}


def _load_template(name: str) -> Dict[str, str]:
    path = os.path.join(_TEMPLATE_DIR, name + ".json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    if name in _PACKAGED:
        return dict(_PACKAGED[name])
    raise ValueError(f"Can't read {path}")


class Prompter:
    __slots__ = ("template", "_verbose")

    def __init__(self, template_name: str = "", verbose: bool = False):
        name = template_name or _DEFAULT_NAME
        self.template = _load_template(name)
        self._verbose = verbose
        if verbose:
            print(f"Using prompt template {name}: {self.template['description']}")

    def generate_prompt(self, instruction: str, input: Optional[str] = None, label: Optional[str] = None) -> str:
        key, fields = ("prompt_input", dict(instruction=instruction, input=input)) if input else \
                      ("prompt_no_input", dict(instruction=instruction))
        text = self.template[key].format(**fields) + (label or "")
        if self._verbose:
            print(text)
        return text

    def get_response(self, output: str) -> str:
        marker = self.template["response_split"]
        return output.rsplit(marker, 1)[-1].strip() if marker in output else output.strip()
