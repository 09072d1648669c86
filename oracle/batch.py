"""Integer page-table metadata, bit-exact restatement.

Follows KVCacheState::cache_slots (xllm/core/framework/request/sequence_kv_state.cpp:86-104)
and BatchInputBuilder::setup_kv_cache_info / padding_decode_batch_size
(xllm/core/framework/batch/batch_input_builder.cpp:739-831, 833-875, 919-937).
Pinned by BatchTest.Basic (tests/core/framework/batch/batch_test.cpp:403-546).
"""
from dataclasses import dataclass, field
from typing import List


@dataclass
class SeqState:
    """One sequence as the scheduler hands it to the batch builder."""
    block_ids: List[int]          # blocks owned by the sequence, in order
    n_kv_cache_tokens: int        # tokens already in the KV cache
    seq_len: int                  # tokens after this step (cached + q_len)

    @property
    def q_len(self) -> int:
        return self.seq_len - self.n_kv_cache_tokens


@dataclass
class PagedMeta:
    new_cache_slots: List[int] = field(default_factory=list)
    paged_kv_indptr: List[int] = field(default_factory=lambda: [0])
    paged_kv_indices: List[int] = field(default_factory=list)
    paged_kv_last_page_len: List[int] = field(default_factory=list)
    q_cu_seq_lens: List[int] = field(default_factory=lambda: [0])
    kv_cu_seq_lens: List[int] = field(default_factory=lambda: [0])
    block_tables: List[List[int]] = field(default_factory=list)
    positions: List[int] = field(default_factory=list)

    def padded_block_tables(self) -> List[int]:
        width = max((len(r) for r in self.block_tables), default=0)
        out = []
        for r in self.block_tables:
            out.extend(r + [0] * (width - len(r)))   # block id 0 = padding block
        return out


def cache_slots(block_ids, block_size, pos_start, pos_end):
    # sequence_kv_state.cpp:96-101
    return [block_ids[i // block_size] * block_size + i % block_size for i in range(pos_start, pos_end)]


def build_paged_meta(seqs: List[SeqState], block_size: int, min_decoding_batch_size: int = 0,
                     num_decoding_tokens: int = 1) -> PagedMeta:
    m = PagedMeta()
    for s in seqs:
        m.new_cache_slots += cache_slots(s.block_ids, block_size, s.n_kv_cache_tokens, s.seq_len)
        m.positions += list(range(s.n_kv_cache_tokens, s.seq_len))
        m.paged_kv_indices += list(s.block_ids)                       # batch_input_builder.cpp:790-796
        m.paged_kv_indptr.append(m.paged_kv_indptr[-1] + len(s.block_ids))
        r = s.seq_len % block_size
        m.paged_kv_last_page_len.append(block_size if r == 0 else r)  # :798-800
        m.q_cu_seq_lens.append(m.q_cu_seq_lens[-1] + s.q_len)
        m.kv_cu_seq_lens.append(m.kv_cu_seq_lens[-1] + s.seq_len)
        m.block_tables.append(list(s.block_ids))
    # padding_decode_batch_size (:833-875): padded rows use slot 0 / block 0 / last_page_len 1
    all_decode = all(s.q_len == num_decoding_tokens and s.n_kv_cache_tokens > 0 for s in seqs)
    if seqs and all_decode:
        for _ in range(len(seqs), min_decoding_batch_size):
            m.new_cache_slots += [0] * num_decoding_tokens
            m.positions += [0] * num_decoding_tokens
            m.q_cu_seq_lens.append(m.q_cu_seq_lens[-1] + num_decoding_tokens)
            m.kv_cu_seq_lens.append(m.kv_cu_seq_lens[-1] + num_decoding_tokens)
            m.block_tables.append([])
            m.paged_kv_indices.append(0)
            m.paged_kv_indptr.append(m.paged_kv_indptr[-1] + 1)
            m.paged_kv_last_page_len.append(1)
    return m


def update_llm_decode_metadata(src: dict, dst: dict, actual_num_tokens: int, padded_num_tokens: int, actual_batch_size: int,
                               actual_indices_size: int) -> dict:
    """CPU restatement of llm_decode_metadata_update_kernel (xllm/core/kernels/cuda/llm_decode_metadata_update.cu:29-62):
    dst is modified in place (numpy / torch int32 arrays) and returned.  Rows beyond the written ranges keep their old
    content (the persistent buffers of the graph executor are larger than the live step)."""
    n, pn, b, ni = actual_num_tokens, padded_num_tokens, actual_batch_size, actual_indices_size
    dst["tokens"][:n] = src["tokens"][:n]
    dst["positions"][:n] = src["positions"][:n]
    dst["new_cache_slots"][:n] = src["new_cache_slots"][:n]
    if pn > n:
        dst["tokens"][n:pn] = 0
        dst["new_cache_slots"][n:pn] = 0
    dst["kv_seq_lens"][:b + 1] = src["kv_seq_lens"][:b + 1]
    dst["paged_kv_indptr"][:b + 1] = src["paged_kv_indptr"][:b + 1]
    dst["kv_seq_lens_delta"][:b] = src["kv_seq_lens"][1:b + 1] - src["kv_seq_lens"][:b]
    dst["paged_kv_last_page_len"][:b] = src["paged_kv_last_page_len"][:b]
    dst["paged_kv_indices"][:ni] = src["paged_kv_indices"][:ni]
    return dst
