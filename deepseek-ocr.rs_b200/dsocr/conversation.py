"""Prompt templates of the reference (crates/core/src/conversation/mod.rs) and `render_prompt`
(crates/core/src/inference.rs:212-225): the CLI renders `--prompt` through `--template` (default `plain`) before it splits
the result on `<image>`.  Four registered templates: deepseek, deepseekv2, plain, alignment."""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

EOS = "<｜end▁of▁sentence｜>"


@dataclass
class ConversationTemplate:
    name: str = ""
    system_template: str = "{system_message}"
    system_message: str = ""
    roles: Tuple[str, str] = ("USER", "ASSISTANT")
    messages: List[Tuple[str, Optional[str]]] = field(default_factory=list)
    sep_style: str = "deepseek"          # deepseek | deepseekv2 | plain | alignment
    sep: str = "\n"
    sep2: Optional[str] = None
    stop_str: List[str] = field(default_factory=list)
    stop_token_ids: List[int] = field(default_factory=list)

    def set_system_message(self, m: str) -> None:
        self.system_message = m

    def append_message(self, role: str, message: Optional[str]) -> None:
        self.messages.append((role, message))

    def update_last_message(self, message: str) -> None:
        if self.messages:
            self.messages[-1] = (self.messages[-1][0], message)

    def reset_messages(self) -> None:
        self.messages = []

    @staticmethod
    def _content(m: Optional[str]) -> Optional[str]:
        if m is None:
            return None
        m = m.strip()
        return m or None

    def get_prompt(self) -> str:
        seps = (self.sep, self.sep2 or "")
        out = ""
        if self.sep_style in ("deepseek", "deepseekv2"):
            system = self.system_template.replace("{system_message}", self.system_message)
            if system:
                out += system + seps[0]
        if self.sep_style == "deepseek":
            for i, (role, m) in enumerate(self.messages):
                c = self._content(m)
                out += f"{role}: {c}{seps[i % 2]}" if c is not None else f"{role}:"
        elif self.sep_style == "deepseekv2":
            for role, m in self.messages:
                c = self._content(m)
                if c is not None:
                    out += ("<｜sft▁begin｜>\n" + c + seps[0]) if role == "User" else (c + seps[1])
        elif self.sep_style == "plain":
            for i, (_, m) in enumerate(self.messages):
                c = self._content(m)
                if c is not None:
                    out += c + seps[i % 2]
        elif self.sep_style == "alignment":
            for i, (_, m) in enumerate(self.messages):
                c = self._content(m)
                if c is not None:
                    out += ("<image>\n" + seps[0]) if i % 2 == 0 else (c + seps[1])
        else:
            raise ValueError(f"unknown separator style {self.sep_style}")
        return out


_TEMPLATES: Dict[str, ConversationTemplate] = {
    "deepseek": ConversationTemplate(name="deepseek", roles=("<|User|>", "<|Assistant|>"), sep_style="deepseek", sep="\n\n", sep2=EOS,
                                     stop_str=["User:", EOS], stop_token_ids=[100001]),
    # registered with SeparatorStyle::DeepSeek in the reference (mod.rs:207-221), kept as is
    "deepseekv2": ConversationTemplate(name="deepseekv2", roles=("<｜User｜>", "<｜Assistant｜>"), sep_style="deepseek", sep="", sep2=EOS,
                                       stop_str=["User:", EOS], stop_token_ids=[100001]),
    "plain": ConversationTemplate(name="plain", system_template="", roles=("", ""), sep_style="plain", sep="", sep2="",
                                  stop_str=["</s>"], stop_token_ids=[100001]),
    "alignment": ConversationTemplate(name="alignment", system_template="", roles=("", ""), sep_style="alignment", sep="", sep2="",
                                      stop_str=["</s>"], stop_token_ids=[100001]),
}


def get_conv_template(name: str) -> Optional[ConversationTemplate]:
    t = _TEMPLATES.get(name)
    return copy.deepcopy(t) if t is not None else None


def render_prompt(template: str, system_prompt: str, raw_prompt: str) -> str:
    t = get_conv_template(template)
    if t is None:
        raise ValueError(f"unknown conversation template {template}")
    t.set_system_message(system_prompt)
    t.reset_messages()
    t.append_message("User", raw_prompt)
    t.append_message("Assistant", None)
    return t.get_prompt()
