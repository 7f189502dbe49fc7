"""Mirrors of the reference's ``tools`` entry points that sit on the hot path:
``worse_only.evalSEA`` and the ``infer.evaluate`` / ``infer.eval_performance`` bookkeeping."""
