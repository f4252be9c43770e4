from .utils import InsertPostInitMethodToModuleSubClasses, call_to_str, substitute_init_recursively

__all__ = ["InsertPostInitMethodToModuleSubClasses", "substitute_init_recursively", "call_to_str"]
