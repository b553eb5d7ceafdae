"""``HasParamsMixin.new_with`` (reference ``src/temgym_core/tree_utils.py:168-181``).

Only the part of the reference's tree utilities the hot path needs; the
``PathBuilder`` machinery of ``run_with_grads`` is out of scope.
"""
import dataclasses


class HasParamsMixin:
    def new_with(self, **kwargs):
        return dataclasses.replace(self, **kwargs)
