"""A minimal in-process stand-in for the ROS surface the reference's nodes touch (test infrastructure).

`src/mav_mppi/scripts/kinova.py` and `drone.py` use: `rospy.init_node`, `rospy.Rate(hz).sleep()`,
`rospy.is_shutdown()`, `rospy.Subscriber(topic, type, callback)`, `rospy.Publisher(topic, type, queue_size=)`
and the message types `sensor_msgs.msg.JointState`, `std_msgs.msg.Float64MultiArray`
(kinova.py:6-10,60-100; drone.py:5-9,86-100).  `install()` registers modules with exactly that surface in
`sys.modules`; messages published on a topic are delivered to the topic's subscribers on ONE background
thread per topic, which is how rospy runs callbacks -- so `update_joint` really races the control step.
"""
from __future__ import annotations

import queue
import sys
import threading
import types


class JointState:
    def __init__(self):
        self.name, self.position, self.velocity, self.effort = [], [], [], []


class Float64MultiArray:
    def __init__(self):
        self.data = []


class _Topic:
    def __init__(self, name, errors):
        self.name = name
        self.errors = errors
        self.callbacks = []
        self.log = []                      # every message ever published (test inspection)
        self.q: "queue.Queue" = queue.Queue()
        self.thread = threading.Thread(target=self._pump, name=f"fake_ros:{name}", daemon=True)
        self.thread.start()

    def _pump(self):
        while True:
            msg = self.q.get()
            if msg is _STOP:
                return
            try:
                for cb in list(self.callbacks):
                    cb(msg)
            except Exception as e:             # a dying callback thread must not hang the test: record and go on
                self.errors.append(repr(e))
            finally:
                self.q.task_done()


_STOP = object()


class Master:
    """Topic registry + the shutdown flag and a virtual clock the test drives."""

    def __init__(self):
        self.topics: dict[str, _Topic] = {}
        self.shutdown = False
        self.errors: list[str] = []
        self.ticks = 0
        self.on_sleep = None               # callable(master) run by Rate.sleep (the simulator hook)
        self.lock = threading.Lock()

    def topic(self, name) -> _Topic:
        with self.lock:
            if name not in self.topics:
                self.topics[name] = _Topic(name, self.errors)
            return self.topics[name]

    def publish(self, name, msg):
        t = self.topic(name)
        t.log.append(msg)
        t.q.put(msg)

    def drain(self):
        for t in list(self.topics.values()):
            t.q.join()

    def close(self):
        self.shutdown = True
        for t in list(self.topics.values()):
            t.q.put(_STOP)


def install(master: Master | None = None) -> Master:
    """Put fake `rospy`, `sensor_msgs.msg`, `std_msgs.msg` into sys.modules; returns the Master."""
    m = master or Master()

    rospy = types.ModuleType("rospy")

    class Rate:
        def __init__(self, hz):
            self.hz = hz

        def sleep(self):
            m.ticks += 1
            if m.on_sleep is not None:
                m.on_sleep(m)

    class Subscriber:
        def __init__(self, topic, msg_type, callback, queue_size=None):
            m.topic(topic).callbacks.append(callback)

    class Publisher:
        def __init__(self, topic, msg_type, queue_size=None):
            self.topic = topic

        def publish(self, msg):
            m.publish(self.topic, msg)

    rospy.init_node = lambda *a, **k: None
    rospy.is_shutdown = lambda: m.shutdown
    rospy.Rate, rospy.Subscriber, rospy.Publisher = Rate, Subscriber, Publisher
    rospy.loginfo = rospy.logwarn = lambda *a, **k: None

    sensor_msgs, sensor_msgs_msg = types.ModuleType("sensor_msgs"), types.ModuleType("sensor_msgs.msg")
    sensor_msgs_msg.JointState = JointState
    sensor_msgs.msg = sensor_msgs_msg
    std_msgs, std_msgs_msg = types.ModuleType("std_msgs"), types.ModuleType("std_msgs.msg")
    std_msgs_msg.Float64MultiArray = Float64MultiArray
    std_msgs.msg = std_msgs_msg

    for name, mod in (("rospy", rospy), ("sensor_msgs", sensor_msgs), ("sensor_msgs.msg", sensor_msgs_msg),
                      ("std_msgs", std_msgs), ("std_msgs.msg", std_msgs_msg)):
        sys.modules[name] = mod
    return m


def uninstall():
    for name in ("rospy", "sensor_msgs", "sensor_msgs.msg", "std_msgs", "std_msgs.msg"):
        sys.modules.pop(name, None)
