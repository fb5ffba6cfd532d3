"""oracle/ -- CPU restatement of the reference's audio-prompt hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker (or as the timed CPU baseline) -- never as the thing shipped or measured as
the product. The product path (llm_speech_summarization_b200/) never imports it and has no CPU fallback.

What is restated (plain torch fp32/fp64 on the CPU, no transformers import, no /root/reference import):
  * the reference's own code on the path: AudioEncoder.forward pool+project (REF/model/audio_encoder.py:56-88),
    utils.py (compute_num_audio_embeds, merge_*, construct_attention_mask, batch_full_embed_sequence,
    soft_cross_entropy; REF/utils.py:13-178), the response-span CE of AudioLlamaForCausalLM.forward
    (REF/model/audio_llama.py:63-101), the loss mixing of the train step (REF/trainer.py:270-374) and the prompt
    assembly of generate_audio_response (REF/inference.py:95-137);
  * the third-party arithmetic those call into, which is NOT vendored under /root/reference: transformers
    (pinned 4.47.0 in REF/requirements.txt:15; 5.5.0 installed here) HubertModel
    (models/hubert/modeling_hubert.py), LlamaModel (models/llama/modeling_llama.py) and the llama3 RoPE scaling
    (modeling_rope_utils.py:550-625), restated from their published algorithms.

Parity pin: the reference has no tests, golden vectors or fixtures of its own, so the pin
is "outputs of the reference itself run here": oracle/make_golden.py imports the reference's modules from
/root/reference (plus the installed transformers) in the build container, runs them on seeded tiny
configurations, checks this restatement against them and commits the resulting vectors under tests/golden/.
tests/test_cpu.py re-checks the restatement against those vectors on every run (CPU, no reference needed); the
`-m gpu` tests then compare the CUDA path with both. The train-mode restatement (dropout sites, LayerDrop,
SpecAugment under explicit masks, oracle/regularizers.py) is pinned the same way against the reference's AudioEncoder
in .train() mode with HF's randomness replaced by the same masks (make_golden.py:run_train_mode_case ->
tests/golden/tiny_hubert_train_mode.pt).
"""
