"""CPU oracle for the WICCA HaarCoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``wicca_b200/`` may import this
package; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the timed CPU baseline, never as the product path.
"""
