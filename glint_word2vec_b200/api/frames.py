"""Tabular-data adapters: the DataFrame side of the Spark ML API without Spark.

The reference's ML layer works on Spark ``Dataset``s (ML:284-305, ML:432-460).
Here a "dataset" may be a pandas ``DataFrame``, a pyarrow ``Table``, a dict of
equal-length columns, or (for ``fit`` only) a plain iterable of token lists.
``transform`` returns the same kind of object it was given with the output
column appended LAST and all other columns kept (SPEC:260-288).
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np


def _is_pandas(data) -> bool:
    try:
        import pandas as pd
        return isinstance(data, pd.DataFrame)
    except ImportError:  # pragma: no cover
        return False


def _is_arrow(data) -> bool:
    try:
        import pyarrow as pa
        return isinstance(data, pa.Table)
    except ImportError:  # pragma: no cover
        return False


def column_names(data) -> Optional[List[str]]:
    if _is_pandas(data):
        return list(data.columns)
    if _is_arrow(data):
        return list(data.column_names)
    if isinstance(data, dict):
        return list(data.keys())
    return None


def num_rows(data) -> int:
    if _is_pandas(data):
        return len(data)
    if _is_arrow(data):
        return data.num_rows
    if isinstance(data, dict):
        return len(next(iter(data.values()))) if data else 0
    return len(data)


def get_column(data, name: str) -> list:
    if _is_pandas(data):
        return data[name].tolist()
    if _is_arrow(data):
        return data.column(name).to_pylist()
    if isinstance(data, dict):
        return list(data[name])
    raise TypeError(f"unsupported dataset type {type(data).__name__}")


def first_non_null(data, name: str):
    if _is_arrow(data):
        col = data.column(name)
        for i in range(min(len(col), 64)):
            v = col[i].as_py()
            if v is not None:
                return v
        return None
    for v in get_column(data, name)[:64] if not _is_pandas(data) else data[name].head(64).tolist():
        if v is not None:
            return v
    return None


def append_column(data, name: str, values: Sequence[Any]):
    """Return a new dataset of the same kind with ``name`` appended last."""
    if _is_pandas(data):
        out = data.copy()
        out[name] = list(values)
        return out
    if _is_arrow(data):
        import pyarrow as pa
        arr = pa.array([None if v is None else np.asarray(v, dtype=np.float64).tolist() for v in values],
                       type=pa.list_(pa.float64()))
        return data.append_column(name, arr)
    if isinstance(data, dict):
        out = dict(data)
        out[name] = list(values)
        return out
    raise TypeError(f"unsupported dataset type {type(data).__name__}")


def make_frame(columns: dict):
    """Build the default tabular result (pandas when available, else dict)."""
    try:
        import pandas as pd
        return pd.DataFrame(columns)
    except ImportError:  # pragma: no cover
        return columns
