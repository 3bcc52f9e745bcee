"""torch-CPU stand-in for the reference's JaxSeq GPT-2 policy (oracle; used by bench.py's CPU baseline and tests only).

Mirrors ``GPT2PPOPolicy.act`` (LLM_RL/algorithms/ppo/gpt2/interface.py:507-546) step by step — join the history into
a prompt, feed finished rows the EOS string, tokenise, LEFT pad / LEFT truncate (BlockingStrategy), sample with
``do_sample=True``, T=1, no top-k/p, EOS = newline id, ``max_new_tokens``; decode, strip the prompt, apply
``out_str_process`` — but calls ``transformers.GPT2LMHeadModel.generate`` on torch-CPU because the JAX stack is not
installable here.  Same weights and tokenizer table as the GPU engine.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from lmrl_gym_b200.environment import BatchedTextPolicy, Text, text_history_to_str


def hf_model_from_params(params, cfg):
    from transformers import GPT2Config as HFConfig, GPT2LMHeadModel
    hf = GPT2LMHeadModel(HFConfig(vocab_size=cfg.n_vocab_rows, n_positions=cfg.n_pos, n_embd=cfg.d_model, n_layer=cfg.n_layer,
                                  n_head=cfg.n_head, n_inner=cfg.d_ff, attn_pdrop=0.0, embd_pdrop=0.0, resid_pdrop=0.0,
                                  layer_norm_epsilon=cfg.ln_eps)).eval()
    sd = {}
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
    sd["transformer.wte.weight"] = t(params["wte"])
    sd["transformer.wpe.weight"] = t(params["wpe"])
    sd["transformer.ln_f.weight"] = t(params["lnf_g"])
    sd["transformer.ln_f.bias"] = t(params["lnf_b"])
    sd["lm_head.weight"] = sd["transformer.wte.weight"]
    names = dict(ln1_g="ln_1.weight", ln1_b="ln_1.bias", attn_w="attn.c_attn.weight", attn_b="attn.c_attn.bias",
                 proj_w="attn.c_proj.weight", proj_b="attn.c_proj.bias", ln2_g="ln_2.weight", ln2_b="ln_2.bias",
                 fc_w="mlp.c_fc.weight", fc_b="mlp.c_fc.bias", fc2_w="mlp.c_proj.weight", fc2_b="mlp.c_proj.bias")
    for l, layer in enumerate(params["layers"]):
        for k, n in names.items():
            sd[f"transformer.h.{l}.{n}"] = t(layer[k])
    hf.load_state_dict(sd, strict=False)
    return hf


class TorchCPUGPT2Policy(BatchedTextPolicy):
    def __init__(self, params, cfg, tokenizer, max_input_length: int = 256, max_new_tokens: int = 6,
                 pad_to_max_input: bool = False, seed: int = 0, out_str_process=None):
        self.model = hf_model_from_params(params, cfg)
        self.cfg, self.tok = cfg, tokenizer
        self.max_input_length, self.max_new_tokens = max_input_length, max_new_tokens
        self.pad_to_max_input = pad_to_max_input          # True = the reference's static shape (eval_bc_gpt2.py:100-104)
        self.out_str_process = out_str_process or (lambda s: s.removesuffix("\n") + "\n")
        self.gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)

    @torch.no_grad()
    def act(self, text_history: List[Optional[tuple]], done: Optional[List[bool]] = None):
        if done is None:
            done = [False] * len(text_history)
        eos, pad = self.tok.newline_id(), self.tok.pad_token_id
        prompts = ["\n" if d else text_history_to_str(h) for h, d in zip(text_history, done)]   # EOS string for done rows
        toks = [self.tok.encode(p)[-self.max_input_length:] for p in prompts]
        width = self.max_input_length if self.pad_to_max_input else max(len(t) for t in toks)
        ids = torch.full((len(toks), width), pad, dtype=torch.long)
        mask = torch.zeros((len(toks), width), dtype=torch.long)
        for i, t in enumerate(toks):
            ids[i, width - len(t):] = torch.tensor(t, dtype=torch.long)
            mask[i, width - len(t):] = 1
        out = self.model.generate(input_ids=ids, attention_mask=mask, do_sample=True, max_new_tokens=self.max_new_tokens,
                                  eos_token_id=eos, pad_token_id=pad, top_k=0, top_p=1.0, temperature=1.0,
                                  suppress_tokens=[pad])
        res = []
        for i, (h, d) in enumerate(zip(text_history, done)):
            if d:
                res.append(None)
                continue
            comp = [t for t in out[i, width:].tolist() if t != pad]
            if eos in comp:
                comp = comp[:comp.index(eos) + 1]
            res.append(tuple(h) + (Text(self.out_str_process(self.tok.decode(comp)), True),))
        return res
