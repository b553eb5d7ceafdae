"""Parameter references (reference ``src/temgym_core/tree_utils.py``).

``obj.params.focal_length`` builds a symbolic reference to a parameter of a component or ray,
the way the reference's ``PathBuilder`` does (tree_utils.py:34-122); ``run_with_grads``
(run.py:182-267) takes a sequence of them.  Only what the accelerated path needs is kept: a
reference is (root object, attribute path); resolving it against a model yields tangent seeds
for the CUDA gradient kernel (``tg_trace_grad_f64``).
"""
import dataclasses


class ParamRef:
    """Symbolic path to a leaf of ``root`` (``PathBuilder``, tree_utils.py:34-122)."""
    __slots__ = ("_pr_root", "_pr_path")

    def __init__(self, root, path=()):
        object.__setattr__(self, "_pr_root", root)
        object.__setattr__(self, "_pr_path", tuple(path))

    def __getattr__(self, name):
        if name.startswith("_pr_") or name.startswith("__"):
            raise AttributeError(name)
        return ParamRef(self._pr_root, self._pr_path + (name,))

    def __getitem__(self, idx):
        return ParamRef(self._pr_root, self._pr_path + (idx,))

    def _resolve_root(self):
        return self._pr_root

    def _resolve(self):
        v = self._pr_root
        for k in self._pr_path:
            v = getattr(v, k) if isinstance(k, str) else v[k]
        return v

    def _build(self, original: bool = True):
        """``(root, key, key, ...)`` -- the dictionary key ``run_with_grads`` reports
        gradients under (tree_utils.py:88-98 with ``original=True``)."""
        return (self._pr_root,) + self._pr_path

    def __repr__(self):
        return f"ParamRef({type(self._pr_root).__name__}{''.join('.' + str(k) for k in self._pr_path)})"


class HasParamsMixin:
    def new_with(self, **kwargs):  # tree_utils.py:169-170
        return dataclasses.replace(self, **kwargs)

    @property
    def params(self):  # tree_utils.py:172-181
        return ParamRef(self)
