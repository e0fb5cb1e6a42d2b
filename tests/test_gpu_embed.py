"""N4: token ids -> T5 embedding gather + image-token concat + mask, against the reference's own torch ops
(/root/reference/architectures/T5VisionModel.py:169-181 restated with torch.nn.functional.embedding / torch.cat)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference(table, image_tokens, input_ids, attention_mask):
    q = torch.nn.functional.embedding(input_ids, table)                       # self.T5_model.shared(input_ids)       :169
    if image_tokens is None:
        return q, attention_mask                                              # "only use question"                  :178-180
    image_mask = torch.ones((image_tokens.shape[0], image_tokens.shape[1]))   #                                      :172
    mask = torch.cat((image_mask, attention_mask.cpu()), axis=1).to(table.device)   #                                :173
    return torch.cat((image_tokens, q), axis=1), mask                         #                                      :176


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("b,L,n_image,hidden", [(16, 37, 50, 512), (3, 5, 0, 512), (128, 64, 50, 768), (1, 1, 1, 8)])
def test_embed_prompt_equals_torch_ops(dtype, b, L, n_image, hidden):
    from multimodalpromptretrieval_b200.embed import embed_prompt
    if dtype != torch.float32 and hidden % 8:
        pytest.skip("rows must be whole 16-byte vectors")
    if dtype == torch.float32 and hidden % 4:
        pytest.skip("rows must be whole 16-byte vectors")
    g = torch.Generator().manual_seed(b * 1000 + L)
    vocab = 3210
    table = torch.randn(vocab, hidden, generator=g).to(dtype).to(DEV)
    stride = L + 11                                                            # ids come as views of a wider block
    ids_block = torch.randint(0, vocab, (b, stride), generator=g).to(DEV)
    lens = torch.randint(1, L + 1, (b,), generator=g)
    mask_block = (torch.arange(stride)[None, :] < lens[:, None]).long().to(DEV)
    ids, mask = ids_block[:, :L], mask_block[:, :L]
    image = torch.randn(b, n_image, hidden, generator=g).to(dtype).to(DEV) if n_image else None
    out, out_mask = embed_prompt(table, ids, mask, image)
    ref, ref_mask = reference(table, image, ids, mask)
    assert out.dtype == ref.dtype and torch.equal(out, ref)                    # a gather: bit-exact
    assert out_mask.dtype == ref_mask.dtype and torch.equal(out_mask, ref_mask)


def test_embed_prompt_backward_matches_autograd():
    from multimodalpromptretrieval_b200.embed import embed_prompt
    g = torch.Generator().manual_seed(5)
    b, L, n_image, hidden, vocab = 8, 12, 50, 512, 300
    table = torch.randn(vocab, hidden, generator=g).to(DEV).requires_grad_()
    image = torch.randn(b, n_image, hidden, generator=g).to(DEV).requires_grad_()
    ids = torch.randint(0, vocab, (b, L), generator=g).to(DEV)
    mask = torch.ones(b, L, dtype=torch.int64, device=DEV)
    w = torch.randn(b, n_image + L, hidden, generator=g).to(DEV)
    out, _ = embed_prompt(table, ids, mask, image)
    (out * w).sum().backward()
    gt, gi = table.grad.clone(), image.grad.clone()
    table.grad = image.grad = None
    ref, _ = reference(table, image, ids, mask)
    (ref * w).sum().backward()
    assert torch.allclose(gt, table.grad, atol=1e-5) and torch.equal(gi, image.grad)


def test_embed_prompt_rejects_out_of_range_ids():
    from multimodalpromptretrieval_b200 import kernels as K
    from multimodalpromptretrieval_b200.embed import embed_prompt
    table = torch.randn(10, 8, device=DEV)
    ids = torch.tensor([[1, 99, 3]], device=DEV)
    out, _ = embed_prompt(table, ids, torch.ones_like(ids))
    assert torch.equal(out[0, 1], torch.zeros(8, device=DEV)) and torch.equal(out[0, 0], table[1])
    assert K.handle(0).device_error() == 301


def test_prepare_input_chain_from_retrieval_to_embeddings(golden_cases, tokenizer):
    """retrieve_prompt_ids (device) -> embed_prompt == the reference's tokenizer -> shared -> cat on the golden prompts."""
    from multimodalpromptretrieval_b200.bank import RetrievalBank
    from multimodalpromptretrieval_b200.embed import embed_prompt
    g = golden_cases["k5_train"]
    table_q = {q: e for q, e in zip(g.questions, g.q_txt)}

    class Clip:
        encode_image = staticmethod(lambda x: x)
        encode_text = staticmethod(lambda x: x)

    bank = RetrievalBank(clip_model=Clip(), clip_tokenize=lambda qs: torch.stack([table_q[q] for q in qs], 0),
                         tokenizer=tokenizer, shard=False)
    bank.install_bank([(g.bank_img, g.bank_txt)], g.answers, g.info, is_training_phase=g.training, retrieval_k=g.k)
    batch = {"image": g.q_img.clone(), "question": g.questions, "task": g.tasks}
    ids, mask = bank.retrieve_prompt_ids(batch, copy=False)
    gen = torch.Generator().manual_seed(1)
    table = torch.randn(len(tokenizer), 512, generator=gen).to(DEV)
    image = torch.randn(g.b, 50, 512, generator=gen).to(DEV)
    out, out_mask = embed_prompt(table, ids, mask, image)
    ref_ids = torch.from_numpy(g.z["input_ids_quant"]).to(DEV)
    ref_mask = torch.from_numpy(g.z["attention_mask_quant"]).to(DEV)
    ref, ref_m = reference(table, image, ref_ids, ref_mask)
    assert torch.equal(out, ref) and torch.equal(out_mask, ref_m)
