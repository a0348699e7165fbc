"""Decode / post-processing half of the evaluation driver (reference evaluation_aqa_dataset.py:85-97, 289-301, 329-384) as
importable functions, so an eval loop around the drop-in `Myriad.generate` produces the reference's result records:

    outputs  = model.generate(samples, **default_generate_kwargs(stopping_criteria))
    records  = postprocess_generate(outputs, model.llama_tokenizer, data_sample, samples, task_type)

The arithmetic here is host-side bookkeeping on <= 90 token ids per sample; the device path ends at the token ids."""
import numpy as np
import torch

STOP_WORD_IDS = ((835,), (2277, 29937))  # '###' in its two LLaMA tokenisations (evaluation_aqa_dataset.py:271-272)


def default_generate_kwargs(stopping_criteria=None):
    """evaluation_aqa_dataset.py:289-301: nucleus sampling with top_p = 0.01 — which keeps only the arg-max token, i.e. greedy."""
    return {"max_new_tokens": 90, "stopping_criteria": stopping_criteria, "do_sample": True, "use_cache": True, "min_length": 1,
            "top_p": 0.01, "temperature": 1.0}


def anomaly_map_handler(output):
    """[B,1,H,W] expert maps in [0,1] -> list of HxWx3 uint8 images (:85-97)"""
    if not isinstance(output, dict) or "ve_anomaly_maps" not in output:
        return []
    maps = output["ve_anomaly_maps"].detach().float().cpu()
    return [(m.expand(3, -1, -1).permute(1, 2, 0).numpy() * 255.0).astype(np.uint8) for m in maps]


def decode_answers(token_ids, tokenizer):
    """ids clamped to [1, 40000] (0 = padding written by stopped rows), batch-decoded, cut at the first '###' (:339-340, :347)"""
    ids = torch.clamp(torch.as_tensor(token_ids), 1, 40000)
    return [t.split("###")[0] for t in tokenizer.batch_decode(ids, add_special_tokens=False)]


def yes_no_error(answer, is_anomaly):
    """'0' when the answer's Yes / No agrees with the ground truth, else '1' (:376-381)"""
    if "Yes" in answer and bool(is_anomaly):
        return "0"
    if "No" in answer and not bool(is_anomaly):
        return "0"
    return "1"


def _item(v):
    return v.item() if hasattr(v, "item") else v


def postprocess_generate(outputs, tokenizer, data_sample, samples, task_type="ad"):
    """One result record per sample, with the reference's keys for each task type (:342-386)."""
    token_ids = outputs["token_ids"] if isinstance(outputs, dict) else outputs
    maps = anomaly_map_handler(outputs)
    answers = decode_answers(token_ids, tokenizer)
    records = []
    for i, text in enumerate(answers):
        question = samples["question"] if len(samples["question"]) == 1 else samples["question"][i]
        if task_type in ("aqa", "roi"):
            rec = {"image_id": _item(data_sample["image_id"][i]), "output": text, "question": question,
                   "options": samples["options"][i].detach().cpu().numpy().tolist()}
            if task_type == "aqa":
                rec["answer"] = _item(data_sample["answer"][i])
            rec["is_anomaly"] = _item(data_sample["is_anomaly"][i])
        elif task_type in ("al", "ad", "adroi", "ad_few", "1cls", "shot"):
            gt = _item(data_sample["is_anomaly"][i])
            rec = {"image_id": _item(data_sample["image_id"][i]), "image_path": "/".join(samples["img_path"][i].split("/")[-5:]),
                   "is_anomaly": gt}
            if maps:
                rec["error"] = yes_no_error(text, gt)
                rec["output"] = text
                rec["anomaly_score"] = str(round(float(maps[i].max()) / 255.0, 4))
        else:
            raise NotImplementedError("Not implement for task type %s" % task_type)
        records.append(rec)
    return records


def summarize(records):
    """accuracy / error counts over records that carry the Yes / No verdict"""
    scored = [r for r in records if "error" in r]
    wrong = sum(r["error"] == "1" for r in scored)
    return {"n": len(scored), "errors": wrong, "accuracy": (1.0 - wrong / len(scored)) if scored else None}
