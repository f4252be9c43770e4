from .utils import (
    DummyDataloader,
    clear_cache_before_run,
    free_port,
    parameterize,
    rerun_if_address_is_in_use,
    rerun_on_exception,
    run_on_environment_flag,
    skip_if_not_enough_gpus,
    spawn,
)
from .comparison import (
    assert_close,
    assert_close_loose,
    assert_equal,
    assert_equal_in_group,
    assert_not_equal,
    check_state_dict_equal,
    assert_hf_output_close,
)

__all__ = [
    "DummyDataloader", "clear_cache_before_run", "free_port", "parameterize", "rerun_if_address_is_in_use",
    "rerun_on_exception", "run_on_environment_flag", "skip_if_not_enough_gpus", "spawn", "assert_close",
    "assert_close_loose", "assert_equal", "assert_equal_in_group", "assert_not_equal", "check_state_dict_equal",
    "assert_hf_output_close",
]
