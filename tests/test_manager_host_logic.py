"""CPU: the EmbeddingManager-shaped facade's whole upload / query / batch_query / similar / delete cycle (the scenario of
tests/test_gpu_manager.py, the reference's call patterns of app/utils/embedder.py:428-617, 784-930) with tests/fake_device.py
standing in for the device: the host half -- ids "doc_<hex>_<item>", metadata, filter pass-through, batch flattening, the LRU
embedding cache and its stats -- is exercised without a GPU; the GPU suite runs the same function against the real kernels."""
from test_collection_host_logic import fake  # noqa: F401  (fixture)
import test_gpu_manager as on_gpu


def test_manager_cycle_on_the_fake_device(fake):  # noqa: F811
    on_gpu.test_manager_upload_query_delete_cycle()


def test_collection_scenarios_of_the_gpu_suite_on_the_fake_device(fake, golden, tmp_path):  # noqa: F811
    """Host-heavy scenarios of the GPU suite -- k beyond the collection and empty collections, ties, type / bitmap filters,
    delete + upsert visibility, the reference's own 70 vectors with their known answers, the Chroma WAL import -- replayed on CPU."""
    import test_gpu_parity as p
    import test_chroma_import as ci
    p.test_k_larger_than_collection_and_empty()
    p.test_duplicates_and_ties_resolve_to_lowest_row()
    p.test_type_filter_and_bitmap_filter()
    p.test_delete_upsert_visibility()
    p.test_golden_fixture_known_answers(golden)
    ci.test_import_reference_wal_on_gpu(tmp_path, golden)
