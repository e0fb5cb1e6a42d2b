"""Prompt token ids -> T5 input embeddings with the image tokens prepended.

Host mirror of the end of ``T5VisionModel.prepare_input`` (/root/reference/architectures/T5VisionModel.py:169-181):

    question_embedding = self.T5_model.shared(encoding["input_ids"].to(self.device))
    image_attn_mask    = torch.ones((image_embeddings.shape[0], image_embeddings.shape[1]))
    attention_mask     = torch.cat((image_attn_mask, encoding.attention_mask), axis=1)
    combined_embedding = torch.cat((image_embeddings, question_embedding), axis=1)      # use_image_info
    combined_embedding = question_embedding; attention_mask = encoding.attention_mask   # only use question

The ids come from the retrieval kernel's tail on the device (``RetrievalBank.retrieve_prompt_ids``), so the forward is
one launch of ``embed_prompt_kernel`` (csrc/embed_gather.cuh) — no H2D copy of ids, no ``torch.cat``.  The backward of
``shared`` (trainable in T5VisionModelFrozen.py:24) and of the image tokens is an index_add / a slice and stays on stock
PyTorch (forward-only kernel, as scoped in SURVEY.md).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import kernels as K


class _EmbedPrompt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, image_tokens, input_ids, attention_mask, mask_dtype):
        n_image = 0 if image_tokens is None else image_tokens.shape[1]
        out, mask = K.embed_prompt(input_ids, attention_mask, table, image_tokens, mask_dtype=mask_dtype)
        ctx.save_for_backward(input_ids)
        ctx.n_image, ctx.table_shape, ctx.has_image = n_image, tuple(table.shape), image_tokens is not None
        ctx.mark_non_differentiable(mask)
        return out, mask

    @staticmethod
    def backward(ctx, grad_out, _grad_mask):
        (input_ids,) = ctx.saved_tensors
        n_image = ctx.n_image
        grad_table = grad_image = None
        if ctx.needs_input_grad[0]:
            grad_table = torch.zeros(ctx.table_shape, dtype=grad_out.dtype, device=grad_out.device)
            grad_table.index_add_(0, input_ids.reshape(-1), grad_out[:, n_image:].reshape(-1, grad_out.shape[-1]))
        if ctx.has_image and ctx.needs_input_grad[1]:
            grad_image = grad_out[:, :n_image]
        return grad_table, grad_image, None, None, None


def embed_prompt(table: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                 image_tokens: Optional[torch.Tensor] = None, mask_dtype: Optional[torch.dtype] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(combined_embedding [B, n_image + L, H], attention_mask [B, n_image + L])`` exactly as ``prepare_input`` builds
    them.  ``table`` = ``T5_model.shared.weight``; ``image_tokens`` = the ``[B, n_image, H]`` CLIP token features (cast to
    the table's dtype if needed) or None ("only use question").  The mask is float32 with image tokens (what the
    reference's ``cat`` of a float ones tensor with the int64 tokenizer mask promotes to), int64 without."""
    if mask_dtype is None:
        mask_dtype = torch.float32 if image_tokens is not None else torch.int64
    if image_tokens is not None:
        if image_tokens.dtype != table.dtype:
            image_tokens = image_tokens.to(table.dtype)
        image_tokens = image_tokens.contiguous()
    ids = input_ids if input_ids.stride(1) == 1 else input_ids.contiguous()
    mask = attention_mask if (attention_mask.stride(1) == 1 and attention_mask.stride(0) == ids.stride(0)) else None
    if mask is None:
        ids, mask = ids.contiguous(), attention_mask.contiguous()
    return _EmbedPrompt.apply(table, image_tokens, ids, mask, mask_dtype)
