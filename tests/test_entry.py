import pytest


def test_build_compiles_all_native_libraries():
    import __graft_entry__ as g

    g.build()


@pytest.mark.gpu
def test_smoke_runs_native_kernels():
    import __graft_entry__ as g

    g.smoke()
