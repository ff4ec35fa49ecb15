"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU checkers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under slam_ros_b200/ does.
"""
