"""B200-native retrieval hot path for MPR_Gen (tossowski/MultimodalPromptRetrieval).

Only the retrieval path lives here: bank build → query-vs-bank scan with fused top-k → candidate merge →
answer vote / prompt-token gather.  Hand-written sm_100a CUDA behind a C ABI (``include/mpr_b200.h``); Python is
the host mirror of the reference's ``VQADataset.create_retrieval_dataset`` / ``retrieve_closest_qa_pairs``.
"""
__version__ = "0.1.0"


def __getattr__(name):   # lazy: importing the package must not require torch/CUDA (build(), CPU tests)
    if name == "RetrievalBank":
        from .bank import RetrievalBank
        return RetrievalBank
    raise AttributeError(name)
