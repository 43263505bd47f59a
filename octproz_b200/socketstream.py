"""Headless mirror of the reference's Socket Stream Extension: the wire format every external consumer of processed OCT data
speaks (octproz_plugins/octproz-socket-stream-extension, docs/docs/plugin-socketstream.md).

  * 13-byte big-endian header  magic 299792458 u32 | payload bytes u32 | frame width u16 | frame height u16 | bit depth u8
    (src/broadcaster.cpp:39,293-304), followed by the raw container bytes of one processed buffer;
  * text protocol on the same socket (src/broadcaster.cpp:262-291): `ping` -> `pong\\n`, `enable_command_only_mode` /
    `disable_command_only_mode` move a connection between the data list and the command list, anything else is handed to the
    host as a remote command (`remote_start`, `set_disp_coeff:...`, ...);
  * transports: TCP/IP, IPC (QLocalServer = a Unix domain socket on Linux) and WebSocket (QWebSocketServer in the reference,
    broadcaster.cpp:74-76,104-110,192-247: every buffer is ONE binary message, header included; commands and replies are text
    messages) -- a minimal RFC 6455 server side here (handshake, unfragmented frames, ping / close), enough for browser clients.
Pure host code (stdlib sockets); the payload is the converted output the pipeline streams to the host
(octb200_register_streaming_buffers + callback, the reference's Gpu2HostNotifier path).
"""
from __future__ import annotations

import base64
import hashlib
import math
import os
import selectors
import socket
import struct
import threading
from dataclasses import dataclass

START_IDENTIFIER = 299792458                      # broadcaster.cpp:39
HEADER_FORMAT = ">IIHHB"                          # QDataStream BigEndian: quint32 quint32 quint16 quint16 quint8 (broadcaster.cpp:295-301)
HEADER_SIZE = struct.calcsize(HEADER_FORMAT)      # 13

MODE_IPC, MODE_TCPIP, MODE_WEBSOCKET = "ipc", "tcpip", "websocket"    # CommunicationMode (socketstreamextensionparameters.h:10-14)
WS_GUID = "258EAFA5-E914-47DA-95CA-C5AB0DC85B11"  # RFC 6455 section 1.3
WS_TEXT, WS_BINARY, WS_CLOSE, WS_PING, WS_PONG = 0x1, 0x2, 0x8, 0x9, 0xA


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    chunks, got = [], 0
    while got < n:
        b = sock.recv(min(1 << 20, n - got))
        if not b:
            raise ConnectionError("socket closed")
        chunks.append(b); got += len(b)
    return b"".join(chunks)


def ws_accept_key(key: str) -> str:
    return base64.b64encode(hashlib.sha1((key + WS_GUID).encode()).digest()).decode()


def ws_frame(opcode: int, payload: bytes, mask: bytes | None = None) -> bytes:
    """one unfragmented frame; servers send unmasked, clients masked (RFC 6455 section 5.2)"""
    n = len(payload)
    head = bytes([0x80 | opcode])
    mbit = 0x80 if mask else 0
    if n < 126:
        head += bytes([mbit | n])
    elif n < (1 << 16):
        head += bytes([mbit | 126]) + struct.pack(">H", n)
    else:
        head += bytes([mbit | 127]) + struct.pack(">Q", n)
    if mask:
        payload = bytes(b ^ mask[i & 3] for i, b in enumerate(payload)) if n < 4096 else _ws_mask_fast(payload, mask)
        return head + mask + payload
    return head + payload


def _ws_mask_fast(payload: bytes, mask: bytes) -> bytes:
    import numpy as np
    a = np.frombuffer(payload, np.uint8)
    m = np.frombuffer((mask * ((len(payload) + 3) // 4))[: len(payload)], np.uint8)
    return (a ^ m).tobytes()


def ws_read_frame(sock: socket.socket):
    """(opcode, payload) of the next frame, unmasked if it was masked"""
    b0, b1 = _recv_exact(sock, 2)
    n = b1 & 0x7F
    if n == 126:
        n = struct.unpack(">H", _recv_exact(sock, 2))[0]
    elif n == 127:
        n = struct.unpack(">Q", _recv_exact(sock, 8))[0]
    mask = _recv_exact(sock, 4) if (b1 & 0x80) else None
    payload = _recv_exact(sock, n) if n else b""
    if mask:
        payload = bytes(b ^ mask[i & 3] for i, b in enumerate(payload)) if n < 4096 else _ws_mask_fast(payload, mask)
    return b0 & 0x0F, payload


def ws_client_connect(host: str, port: int, timeout: float = 5.0) -> socket.socket:
    """client side of the opening handshake (tests, headless consumers)"""
    c = socket.create_connection((host, port), timeout=timeout)
    key = base64.b64encode(os.urandom(16)).decode()
    c.sendall((f"GET / HTTP/1.1\r\nHost: {host}:{port}\r\nUpgrade: websocket\r\nConnection: Upgrade\r\n"
               f"Sec-WebSocket-Key: {key}\r\nSec-WebSocket-Version: 13\r\n\r\n").encode())
    resp = b""
    while b"\r\n\r\n" not in resp:
        b = c.recv(4096)
        if not b:
            raise ConnectionError("handshake failed")
        resp += b
    if b" 101 " not in resp.split(b"\r\n", 1)[0] or ws_accept_key(key).encode() not in resp:
        raise ConnectionError("bad handshake response")
    return c


def pack_header(buffer_size_in_bytes: int, frame_width: int, frame_height: int, bit_depth: int) -> bytes:
    """the truncating casts of socketstreamextension.cpp:283-289 included (quint32 / quint16 / quint8)"""
    return struct.pack(HEADER_FORMAT, START_IDENTIFIER, buffer_size_in_bytes & 0xFFFFFFFF, frame_width & 0xFFFF,
                       frame_height & 0xFFFF, bit_depth & 0xFF)


def unpack_header(b: bytes) -> dict:
    magic, size, w, h, bits = struct.unpack(HEADER_FORMAT, b[:HEADER_SIZE])
    if magic != START_IDENTIFIER:
        raise ValueError(f"bad magic number {magic}")
    return {"size": size, "width": w, "height": h, "bitDepth": bits}


@dataclass
class SocketStreamExtensionParameters:
    """socketstreamextensionparameters.h:16-23"""
    mode: str = MODE_TCPIP
    pipeName: str = "octproz"
    ip: str = "127.0.0.1"
    port: int = 1234
    sendHeader: bool = True
    autoConnect: bool = False


class Broadcaster:
    """src/broadcaster.cpp, without Qt: one listener thread accepts clients and serves the text protocol; broadcast() writes a
    buffer to every connection that is not in command-only mode."""

    def __init__(self, params: SocketStreamExtensionParameters | None = None, on_remote_command=None, on_info=None):
        self.params = params or SocketStreamExtensionParameters()
        self.on_remote_command, self.on_info = on_remote_command, on_info
        self.isBroadcasting = False
        self._srv = None
        self._sel = None
        self._thread = None
        self._lock = threading.Lock()
        self.dataConnections: list[socket.socket] = []
        self.commandConnections: list[socket.socket] = []
        self._unix_path = None
        self._ws = set()              # connections that completed the WebSocket opening handshake

    def setParams(self, params: SocketStreamExtensionParameters) -> None:
        self.params = params

    # ---- lifecycle (broadcaster.cpp:83-164) ----
    def startBroadcasting(self) -> None:
        self.stopBroadcasting()
        p = self.params
        if p.mode == MODE_TCPIP:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((p.ip, int(p.port)))
        elif p.mode == MODE_IPC:
            path = p.pipeName if os.path.isabs(p.pipeName) else os.path.join("/tmp", p.pipeName)
            if os.path.exists(path):
                os.unlink(path)
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(path)
            self._unix_path = path
        elif p.mode == MODE_WEBSOCKET:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)                       # QHostAddress::Any (broadcaster.cpp:106)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind(("0.0.0.0", int(p.port)))
        else:
            raise ValueError("unknown communication mode")
        srv.listen(16)
        srv.setblocking(False)
        self._srv = srv
        self._sel = selectors.DefaultSelector()
        self._sel.register(srv, selectors.EVENT_READ, "accept")
        self.isBroadcasting = True
        self._thread = threading.Thread(target=self._serve, daemon=True)
        self._thread.start()

    @property
    def address(self):
        return self._srv.getsockname() if self._srv else None

    def stopBroadcasting(self) -> None:
        if not self.isBroadcasting:
            return
        self.isBroadcasting = False
        if self._thread:
            self._thread.join(2.0)
        with self._lock:
            for c in self.commandConnections + self.dataConnections:
                try:
                    c.close()
                except OSError:
                    pass
            self.commandConnections.clear(); self.dataConnections.clear()
        if self._sel:
            self._sel.close(); self._sel = None
        if self._srv:
            self._srv.close(); self._srv = None
        if self._unix_path and os.path.exists(self._unix_path):
            os.unlink(self._unix_path)
        self._unix_path = None

    # ---- listener thread: accept + text protocol (broadcaster.cpp:170-291) ----
    def _serve(self) -> None:
        while self.isBroadcasting:
            try:
                events = self._sel.select(timeout=0.05)
            except (OSError, ValueError):
                return
            for key, _ in events:
                if key.data == "accept":
                    try:
                        conn, _ = key.fileobj.accept()
                    except OSError:
                        continue
                    conn.setblocking(True)
                    if self.params.mode == MODE_WEBSOCKET and not self._ws_handshake(conn):
                        conn.close()
                        continue
                    with self._lock:
                        self.dataConnections.append(conn)           # new clients receive data until they opt out (:186)
                    self._sel.register(conn, selectors.EVENT_READ, "client")
                    if self.on_info:
                        self.on_info("Client connected!")
                else:
                    conn = key.fileobj
                    if conn in self._ws:
                        try:
                            op, data = ws_read_frame(conn)
                        except (OSError, ConnectionError, struct.error):
                            op, data = WS_CLOSE, b""
                        if op == WS_PING:
                            self._send(conn, data, WS_PONG)
                            continue
                        if op == WS_CLOSE:
                            self._drop(conn)
                            continue
                        if op not in (WS_TEXT, WS_BINARY):
                            continue
                    else:
                        try:
                            data = conn.recv(65536)
                        except OSError:
                            data = b""
                        if not data:
                            self._drop(conn)
                            continue
                    self.processIncomingMessage(data.decode("utf-8", "replace").strip(), conn)

    def _ws_handshake(self, conn) -> bool:
        """server side of the opening handshake (RFC 6455 section 4.2)"""
        try:
            conn.settimeout(2.0)
            req = b""
            while b"\r\n\r\n" not in req and len(req) < 16384:
                b = conn.recv(4096)
                if not b:
                    return False
                req += b
            key = None
            for line in req.split(b"\r\n"):
                if line.lower().startswith(b"sec-websocket-key:"):
                    key = line.split(b":", 1)[1].strip().decode()
            if key is None:
                conn.sendall(b"HTTP/1.1 400 Bad Request\r\n\r\n")
                return False
            conn.sendall(("HTTP/1.1 101 Switching Protocols\r\nUpgrade: websocket\r\nConnection: Upgrade\r\n"
                          f"Sec-WebSocket-Accept: {ws_accept_key(key)}\r\n\r\n").encode())
            conn.settimeout(None)
            self._ws.add(conn)
            return True
        except OSError:
            return False

    def _send(self, conn, data: bytes, ws_opcode: int = WS_TEXT) -> None:
        """a reply on the connection's own framing: raw bytes, or one WebSocket message"""
        conn.sendall(ws_frame(ws_opcode, data) if conn in self._ws else data)

    def _drop(self, conn) -> None:
        self._ws.discard(conn)
        try:
            self._sel.unregister(conn)
        except (KeyError, ValueError):
            pass
        with self._lock:
            if conn in self.dataConnections:
                self.dataConnections.remove(conn)
            if conn in self.commandConnections:
                self.commandConnections.remove(conn)
        try:
            conn.close()
        except OSError:
            pass

    def processIncomingMessage(self, dataString: str, device) -> None:
        """broadcaster.cpp:262-291"""
        if dataString == "ping":
            self._send(device, b"pong\n")
        elif dataString == "enable_command_only_mode":
            with self._lock:
                moved = device in self.dataConnections
                if moved:
                    self.dataConnections.remove(device); self.commandConnections.append(device)
            if moved:
                self._send(device, b"Command mode enabled.\n")
        elif dataString == "disable_command_only_mode":
            with self._lock:
                moved = device in self.commandConnections
                if moved:
                    self.commandConnections.remove(device); self.dataConnections.append(device)
            if moved:
                self._send(device, b"Command mode disabled.\n")
        elif self.on_remote_command:
            self.on_remote_command(dataString)

    # ---- data path (broadcaster.cpp:293-325) ----
    def broadcast(self, buffer, bufferSizeInBytes: int, framesPerBuffer: int, frameWidth: int, frameHeight: int, bitDepth: int) -> int:
        payload = memoryview(buffer).cast("B")[:bufferSizeInBytes]
        head = pack_header(bufferSizeInBytes, frameWidth, frameHeight, bitDepth) if self.params.sendHeader else b""
        with self._lock:
            targets = list(self.dataConnections)
        sent = 0
        for c in targets:
            try:
                if c in self._ws:
                    c.sendall(ws_frame(WS_BINARY, head + bytes(payload)))     # one binary message per buffer (broadcaster.cpp:321-325)
                else:
                    c.sendall(head); c.sendall(payload)
                sent += 1
            except OSError:
                self._drop(c)
        return sent


class SocketStreamExtension:
    """the Extension side (src/socketstreamextension.cpp:271-300): turns a processed buffer into one broadcast"""

    def __init__(self, broadcaster: Broadcaster):
        self.broadcastServer = broadcaster
        self.active = True

    def processedDataReceived(self, buffer, bitDepth: int, samplesPerLine: int, linesPerFrame: int, framesPerBuffer: int,
                              buffersPerVolume: int = 1, currentBufferNr: int = 0) -> int:
        if not self.active:
            return 0
        bytes_per_sample = math.ceil(bitDepth / 8.0)
        size = samplesPerLine * linesPerFrame * framesPerBuffer * bytes_per_sample
        return self.broadcastServer.broadcast(buffer, size & 0xFFFFFFFF, framesPerBuffer & 0xFFFF, samplesPerLine, linesPerFrame, bitDepth)


def read_frame(sock: socket.socket, with_header: bool = True, payload_bytes: int | None = None):
    """client side, as in the reference's examples/octproz_tcpip_connection_opencv.py: returns (header dict or None, payload bytes)"""
    def exact(n):
        chunks, got = [], 0
        while got < n:
            b = sock.recv(min(1 << 20, n - got))
            if not b:
                raise ConnectionError("socket closed")
            chunks.append(b); got += len(b)
        return b"".join(chunks)
    if with_header:
        h = unpack_header(exact(HEADER_SIZE))
        return h, exact(h["size"])
    return None, exact(payload_bytes)
