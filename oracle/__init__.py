"""Oracle: CPU restatements of the reference's modal hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package. The product package (mesheditor_b200) never does.
"""
