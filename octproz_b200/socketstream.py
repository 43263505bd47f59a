"""Headless mirror of the reference's Socket Stream Extension: the wire format every external consumer of processed OCT data
speaks (octproz_plugins/octproz-socket-stream-extension, docs/docs/plugin-socketstream.md).

  * 13-byte big-endian header  magic 299792458 u32 | payload bytes u32 | frame width u16 | frame height u16 | bit depth u8
    (src/broadcaster.cpp:39,293-304), followed by the raw container bytes of one processed buffer;
  * text protocol on the same socket (src/broadcaster.cpp:262-291): `ping` -> `pong\\n`, `enable_command_only_mode` /
    `disable_command_only_mode` move a connection between the data list and the command list, anything else is handed to the
    host as a remote command (`remote_start`, `set_disp_coeff:...`, ...);
  * transports: TCP/IP and IPC (QLocalServer = a Unix domain socket on Linux).  The WebSocket mode of the reference needs a
    WebSocket stack and is not provided here.
Pure host code (stdlib sockets); the payload is the converted output the pipeline streams to the host
(octb200_register_streaming_buffers + callback, the reference's Gpu2HostNotifier path).
"""
from __future__ import annotations

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

MODE_IPC, MODE_TCPIP = "ipc", "tcpip"             # CommunicationMode (socketstreamextensionparameters.h:10-14)


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
        else:
            raise ValueError("unknown communication mode (WebSocket mode is not provided)")
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
                    with self._lock:
                        self.dataConnections.append(conn)           # new clients receive data until they opt out (:186)
                    self._sel.register(conn, selectors.EVENT_READ, "client")
                    if self.on_info:
                        self.on_info("Client connected!")
                else:
                    conn = key.fileobj
                    try:
                        data = conn.recv(65536)
                    except OSError:
                        data = b""
                    if not data:
                        self._drop(conn)
                        continue
                    self.processIncomingMessage(data.decode("utf-8", "replace").strip(), conn)

    def _drop(self, conn) -> None:
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
            device.sendall(b"pong\n")
        elif dataString == "enable_command_only_mode":
            with self._lock:
                moved = device in self.dataConnections
                if moved:
                    self.dataConnections.remove(device); self.commandConnections.append(device)
            if moved:
                device.sendall(b"Command mode enabled.\n")
        elif dataString == "disable_command_only_mode":
            with self._lock:
                moved = device in self.commandConnections
                if moved:
                    self.commandConnections.remove(device); self.dataConnections.append(device)
            if moved:
                device.sendall(b"Command mode disabled.\n")
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
                c.sendall(head); c.sendall(payload); sent += 1
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
