"""CPU oracle for the GPIS active-set-selection + prediction path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gpis_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or
the reported CPU baseline -- never as the product path.
"""
