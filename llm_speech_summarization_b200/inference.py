"""Drop-in for REF/inference.py: `LLMSpeechTextInference(config, audio_encoder_checkpoint, device)` with
`generate_audio_response(audio, additional_text_prompt="", max_new_tokens=256) -> str` (REF/inference.py:95-137) and
`generate_text_response` (:76-93), running the encoder and the prompt prefill on the sm_100a path.

Differences that are forced by the environment, not by design: tokenizer files and LLM weights are not reachable
offline, so both can be injected (`tokenizer=`, `llm=`); when they are not given the constructor does exactly what
the reference does (AutoTokenizer / from_pretrained). The waveform is fed as fp32 (the reference's `.half()` cast,
REF/inference.py:97, only exists to match fp16 autocast).
"""
from __future__ import annotations

import torch

from .model.audio_encoder import AudioEncoder
from .model.audio_llama import AudioLlamaForCausalLM
from .utils import (LLAMA_PROMPT_PREFIX, LLAMA_PROMPT_SUFFIX, MINICHAT_PROMPT_PREFIX, MINICHAT_PROMPT_SUFFIX,
                    merge_prompt_tokens)


class LLMSpeechTextInference():
    def __init__(self, config, audio_encoder_checkpoint, device, tokenizer=None, llm=None):
        self.config = config
        self.device = device

        # Audio encoder: the inference checkpoint is a BARE AudioEncoder state_dict (REF/inference.py:24-26);
        # an already-loaded dict is accepted as well as a path.
        checkpoint = audio_encoder_checkpoint
        if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__"):
            checkpoint = torch.load(checkpoint, map_location="cpu")
        if "audio_encoder" in checkpoint and "optimizer" in checkpoint:  # a trainer checkpoint (REF/trainer.py:518-526)
            checkpoint = checkpoint["audio_encoder"]
        self.audio_encoder = AudioEncoder(self.config, self.device)
        self.audio_encoder.load_state_dict(checkpoint)
        self.audio_encoder.eval().to(self.device)

        self.llm_type = self.config.model.llm_type
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(self.llm_type, use_fast=False, padding_side="left")
            tokenizer.pad_token = tokenizer.eos_token
        self.llm_tokenizer = tokenizer

        if "llama" in self.llm_type:
            self.prompt_prefix = LLAMA_PROMPT_PREFIX
            self.prompt_suffix = LLAMA_PROMPT_SUFFIX
        else:
            self.prompt_prefix = MINICHAT_PROMPT_PREFIX
            self.prompt_suffix = MINICHAT_PROMPT_SUFFIX

        if llm is None:
            llm = AudioLlamaForCausalLM.from_pretrained(self.llm_type, use_cache=True, torch_dtype=torch.float16)
        self.llm = llm.eval()
        self.llm.to(self.device)

    def generate_llm_response(self, inputs_embeds, max_new_tokens=256):
        with torch.no_grad():
            generate_ids = self.llm.generate(input_ids=None, inputs_embeds=inputs_embeds,
                                             max_new_tokens=max_new_tokens)  # greedy, like REF/inference.py:60-66
        return self.llm_tokenizer.batch_decode(generate_ids, skip_special_tokens=True,
                                               clean_up_tokenization_spaces=True)

    def generate_text_response(self, input_text, max_new_tokens=256):
        full_text_prompt = f"{self.prompt_prefix} {input_text}{self.prompt_suffix} "
        with torch.no_grad():
            prompt_input_ids = self.llm_tokenizer(full_text_prompt, return_tensors='pt').input_ids.to(self.device)
            prompt_embeds = self.llm.model.embed_tokens(prompt_input_ids)
            llm_response = self.generate_llm_response(inputs_embeds=prompt_embeds, max_new_tokens=max_new_tokens)[0]
        return llm_response

    def prompt_embeddings(self, audio, additional_text_prompt=""):
        """Encoder + prompt assembly of generate_audio_response (REF/inference.py:96-134), returned for inspection."""
        with torch.no_grad():
            audio_tensor = torch.as_tensor(audio, dtype=torch.float32).unsqueeze(0).to(self.device)
            if self.audio_encoder.downsample_method == "ctc_pool":
                # REF/inference.py:100-105 calls self.get_ctc_pool_ranges, which does not exist in the reference
                raise AttributeError("'LLMSpeechTextInference' object has no attribute 'get_ctc_pool_ranges'")
            audio_embeds = self.audio_encoder(audio_tensor, ctc_pool_ranges=None)
            if len(additional_text_prompt) > 0:
                # text prompt first, start-of-sentence token removed (REF/inference.py:116-122)
                ids = self.llm_tokenizer(additional_text_prompt, return_tensors='pt').input_ids[:, 1:].to(self.device)
                text_embeds = self.llm.model.embed_tokens(ids)
                combined_embeds = torch.cat([text_embeds, audio_embeds], dim=1)
            else:
                combined_embeds = audio_embeds
            return merge_prompt_tokens(inputs_embeds=combined_embeds, tokenizer=self.llm_tokenizer,
                                       embed_tokens=self.llm.model.embed_tokens, llm_type=self.llm_type,
                                       device=self.device)

    def generate_audio_response(self, audio, additional_text_prompt="", max_new_tokens=256):
        prompt_emb_sequence = self.prompt_embeddings(audio, additional_text_prompt)
        return self.generate_llm_response(prompt_emb_sequence, max_new_tokens)[0]
