"""No-op replacement for the SIBR viewer hook (reference gaussian_renderer/network_gui.py): the training
scripts call init() once and poll try_connect()/conn every iteration; without a viewer nothing happens."""
host = "127.0.0.1"
port = 6009
conn = None
addr = None


def init(wish_host, wish_port):
    global host, port
    host, port = wish_host, wish_port


def try_connect():
    return None


def receive():
    return None, False, False, False, False, 1.0


def send(message_bytes, verify):
    return None
