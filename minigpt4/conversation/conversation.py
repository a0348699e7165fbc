"""Conversation surface of the reference (minigpt4/conversation/conversation.py:23-229): prompt assembly with the
'###' separator, the row-0 stop criterion, and `Chat` (ask / answer / upload_img / get_context_emb) over the CUDA
model. `Chat.upload_img` follows the CURRENT Myriad signatures (prepare_sample(samples, stage), encode_img(image, maps,
stage)); the reference's Chat still calls the older ones (conversation.py:207-208, stale vs myriad.py:241,313)."""
import dataclasses
from enum import Enum, auto
from typing import Any, List

import torch


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()


@dataclasses.dataclass
class Conversation:
    system: str
    roles: List[str]
    messages: List[List[str]]
    offset: int
    sep_style: SeparatorStyle = SeparatorStyle.SINGLE
    sep: str = "###"
    sep2: str = None
    skip_next: bool = False
    conv_id: Any = None

    def get_prompt(self):
        if self.sep_style == SeparatorStyle.SINGLE:
            out = self.system + self.sep
            for role, msg in self.messages:
                out += (role + ": " + msg + self.sep) if msg else (role + ":")
            return out
        seps = [self.sep, self.sep2]
        out = self.system + seps[0]
        for i, (role, msg) in enumerate(self.messages):
            out += (role + ": " + msg + seps[i % 2]) if msg else (role + ":")
        return out

    def append_message(self, role, message):
        self.messages.append([role, message])

    def to_gradio_chatbot(self):
        out = []
        for i, (role, msg) in enumerate(self.messages[self.offset:]):
            if i % 2 == 0:
                out.append([msg, None])
            else:
                out[-1][-1] = msg
        return out

    def copy(self):
        return Conversation(self.system, self.roles, [[r, m] for r, m in self.messages], self.offset, self.sep_style, self.sep,
                            self.sep2, conv_id=self.conv_id)

    def dict(self):
        return {"system": self.system, "roles": self.roles, "messages": self.messages, "offset": self.offset, "sep": self.sep,
                "sep2": self.sep2, "conv_id": self.conv_id}


class StoppingCriteriaSub:
    """Stops when ROW 0 of input_ids ends with any of `stops` (reference conversation.py:96-107). The device greedy loop
    (myr_greedy_step) evaluates exactly this; the callable form is kept for callers that drive their own loop."""

    def __init__(self, stops=(), encounters=1):
        self.stops = list(stops)

    def __call__(self, input_ids, scores=None, **kwargs):
        for stop in self.stops:
            stop = torch.as_tensor(stop).to(input_ids.device)
            if input_ids.shape[1] >= len(stop) and torch.all(stop == input_ids[0][-len(stop):]).item():
                return True
        return False


class StoppingCriteriaList(list):
    def __call__(self, input_ids, scores=None, **kw):
        return any(c(input_ids, scores, **kw) for c in self)


CONV_VISION = Conversation(
    system="Give the following image: <Img>ImageContent</Img>. You will be able to see the image once I provide it to you. "
           "Please answer my questions.",
    roles=("Human", "Assistant"), messages=[], offset=2, sep_style=SeparatorStyle.SINGLE, sep="###")


class Chat:
    def __init__(self, model, vis_processor=None, device="cuda:0"):
        self.device, self.model, self.vis_processor = device, model, vis_processor
        self.stopping_criteria = StoppingCriteriaList([StoppingCriteriaSub(stops=[torch.tensor([835]), torch.tensor([2277, 29937])])])

    def ask(self, text, conv):
        if conv.messages and conv.messages[-1][0] == conv.roles[0] and conv.messages[-1][1][-6:] == "</Img>":
            conv.messages[-1][1] = " ".join([conv.messages[-1][1], text])
        else:
            conv.append_message(conv.roles[0], text)

    def get_context_emb(self, conv, img_list):
        segs = conv.get_prompt().split("<ImageHere>")
        assert len(segs) == len(img_list) + 1, "Unmatched numbers of image placeholders and images."
        embs = []
        for i, seg in enumerate(segs):
            ids = self.model.llama_tokenizer(seg, return_tensors="pt", add_special_tokens=i == 0).input_ids
            embs.append(self.model.llama_model.model.embed_tokens(ids.to(self.device)))
            if i < len(img_list):
                embs.append(img_list[i])
        return torch.cat(embs, dim=1)

    def answer(self, conv, img_list, max_new_tokens=300, max_length=2000, **kw):
        conv.append_message(conv.roles[1], None)
        embs = self.get_context_emb(conv, img_list)
        begin = max(0, embs.shape[1] + max_new_tokens - max_length)
        out = self.model.llama_model.generate(inputs_embeds=embs[:, begin:], max_new_tokens=max_new_tokens,
                                              stopping_criteria=self.stopping_criteria, min_length=1)
        ids = out[0]
        if len(ids) and int(ids[0]) in (0, 1):
            ids = ids[1:]
        text = self.model.llama_tokenizer.decode(ids, add_special_tokens=False)
        text = text.split("###")[0].split("Assistant:")[-1].strip()
        conv.messages[-1][1] = text
        return text, ids.cpu().numpy()

    def upload_img(self, samples, conv, img_list, stage=1):
        image, _, _, maps, _ = self.model.prepare_sample(samples, stage)
        emb, _ = self.model.encode_img(image.to(self.device), maps.to(self.device), stage)
        img_list.append(emb)
        conv.append_message(conv.roles[0], "<Img><ImageHere></Img>")
        return "Received."
