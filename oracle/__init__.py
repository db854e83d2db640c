"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference rasterizer (oracle/raster_oracle.c, loaded by oracle/cpu.py) and a
loader for the unmodified reference compiled from /root/reference (oracle/refdgr.py, built by
oracle/build_ref.py into oracle/_ref/).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may import this
package -- as the checker, never as the thing measured or shipped.  frosting_b200/ never imports it.
"""
