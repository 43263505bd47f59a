"""Socket Stream wire format and protocol (octproz_b200/socketstream.py) on loopback: header bytes, TCP and IPC transports,
data vs command-only connections, ping, remote commands."""
import socket
import struct
import time

import numpy as np
import pytest

from octproz_b200.socketstream import (HEADER_SIZE, MODE_IPC, MODE_TCPIP, START_IDENTIFIER, Broadcaster, SocketStreamExtension,
                                       SocketStreamExtensionParameters, pack_header, read_frame, unpack_header)


def test_header_bytes_are_the_documented_ones():
    # docs/docs/plugin-socketstream.md "Data header": 4 + 4 + 2 + 2 + 1 bytes, big endian, magic 299792458
    h = pack_header(512 * 256 * 2, 512, 256, 12)
    assert HEADER_SIZE == 13 and len(h) == 13
    assert h == bytes.fromhex("11de784a") + bytes.fromhex("00040000") + bytes.fromhex("0200") + bytes.fromhex("0100") + b"\x0c"
    assert struct.unpack(">I", h[:4])[0] == START_IDENTIFIER == 299792458
    assert unpack_header(h) == {"size": 262144, "width": 512, "height": 256, "bitDepth": 12}
    # quint16 / quint8 truncation of the reference's casts (socketstreamextension.cpp:283-289)
    assert unpack_header(pack_header(2 ** 32 + 5, 65536 + 7, 3, 256 + 8)) == {"size": 5, "width": 7, "height": 3, "bitDepth": 8}
    with pytest.raises(ValueError):
        unpack_header(b"\x00" * 13)


def _wait(cond, t=2.0):
    end = time.time() + t
    while time.time() < end:
        if cond():
            return True
        time.sleep(0.01)
    return False


@pytest.mark.parametrize("mode", [MODE_TCPIP, MODE_IPC])
def test_broadcast_and_text_protocol(mode, tmp_path):
    cmds = []
    prm = SocketStreamExtensionParameters(mode=mode, ip="127.0.0.1", port=0, pipeName=str(tmp_path / "octproz_pipe"))
    b = Broadcaster(prm, on_remote_command=cmds.append)
    b.startBroadcasting()
    try:
        def client():
            if mode == MODE_TCPIP:
                s = socket.create_connection(b.address, timeout=5)
            else:
                s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM); s.settimeout(5); s.connect(b.address)
            return s
        data, cmd = client(), client()
        assert _wait(lambda: len(b.dataConnections) == 2)
        cmd.sendall(b"ping\n"); assert cmd.recv(64) == b"pong\n"
        cmd.sendall(b"enable_command_only_mode\n"); assert cmd.recv(64) == b"Command mode enabled.\n"
        assert _wait(lambda: len(b.dataConnections) == 1 and len(b.commandConnections) == 1)
        cmd.sendall(b"set_disp_coeff:0:97:-96.6:nullptr\n")
        assert _wait(lambda: cmds == ["set_disp_coeff:0:97:-96.6:nullptr"])
        # one processed buffer: 2 frames of 16 A-scans x 32 samples, 12-bit in u16 containers
        frames = (np.arange(2 * 16 * 32, dtype=np.uint16) * 3 % 4096).reshape(2, 16, 32)
        ext = SocketStreamExtension(b)
        assert ext.processedDataReceived(frames, 12, 32, 16, 2) == 1         # only the data connection receives it
        h, payload = read_frame(data)
        assert h == {"size": frames.nbytes, "width": 32, "height": 16, "bitDepth": 12}
        assert np.array_equal(np.frombuffer(payload, np.uint16).reshape(frames.shape), frames)
        cmd.settimeout(0.2)
        with pytest.raises((socket.timeout, TimeoutError)):
            cmd.recv(1)                                                      # command-only connections get no image data
        # headerless mode (sendHeader = false): payload only
        prm.sendHeader = False
        ext.processedDataReceived(frames, 12, 32, 16, 2)
        _, payload = read_frame(data, with_header=False, payload_bytes=frames.nbytes)
        assert payload == frames.tobytes()
        cmd.settimeout(5)
        cmd.sendall(b"disable_command_only_mode\n"); assert cmd.recv(64) == b"Command mode disabled.\n"
        assert _wait(lambda: len(b.dataConnections) == 2)
        data.close()
        assert _wait(lambda: len(b.dataConnections) == 1)                    # disconnects are noticed
        cmd.close()
    finally:
        b.stopBroadcasting()
    assert not b.isBroadcasting and b.address is None


def test_websocket_mode_binary_messages_and_text_protocol():
    """CommunicationMode::WebSocket (broadcaster.cpp:74-76,192-247,321-325): one binary message per processed buffer (13-byte header
    included), commands and replies as text messages, over an RFC 6455 connection"""
    from octproz_b200.socketstream import (MODE_WEBSOCKET, WS_BINARY, WS_CLOSE, WS_PING, WS_PONG, WS_TEXT, unpack_header, ws_accept_key,
                                           ws_client_connect, ws_frame, ws_read_frame, HEADER_SIZE)
    assert ws_accept_key("dGhlIHNhbXBsZSBub25jZQ==") == "s3pPLMBiTxaQ9kYGzzhZRbK+xOo="      # the known answer of RFC 6455 section 1.3
    cmds = []
    b = Broadcaster(SocketStreamExtensionParameters(mode=MODE_WEBSOCKET, port=0), on_remote_command=cmds.append)
    b.startBroadcasting()
    try:
        port = b.address[1]
        data, cmd = ws_client_connect("127.0.0.1", port), ws_client_connect("127.0.0.1", port)
        assert _wait(lambda: len(b.dataConnections) == 2)
        mask = b"\x11\x22\x33\x44"
        cmd.sendall(ws_frame(WS_TEXT, b"ping", mask)); assert ws_read_frame(cmd) == (WS_TEXT, b"pong\n")
        cmd.sendall(ws_frame(WS_PING, b"abc", mask)); assert ws_read_frame(cmd) == (WS_PONG, b"abc")
        cmd.sendall(ws_frame(WS_TEXT, b"enable_command_only_mode", mask)); assert ws_read_frame(cmd) == (WS_TEXT, b"Command mode enabled.\n")
        assert _wait(lambda: len(b.dataConnections) == 1 and len(b.commandConnections) == 1)
        cmd.sendall(ws_frame(WS_TEXT, b"remote_start", mask))
        assert _wait(lambda: cmds == ["remote_start"])
        # a buffer larger than 64 KiB: 64-bit payload length
        frames = (np.arange(2 * 128 * 256, dtype=np.uint32) * 7 % 4096).astype(np.uint16).reshape(2, 128, 256)
        assert SocketStreamExtension(b).processedDataReceived(frames, 12, 256, 128, 2) == 1
        op, msg = ws_read_frame(data)
        assert op == WS_BINARY and unpack_header(msg) == {"size": frames.nbytes, "width": 256, "height": 128, "bitDepth": 12}
        assert msg[HEADER_SIZE:] == frames.tobytes()
        data.sendall(ws_frame(WS_CLOSE, b"", mask))
        assert _wait(lambda: len(b.dataConnections) == 0)
        cmd.close()
    finally:
        b.stopBroadcasting()
