"""Stand-in for pyglet, just enough for pybfm to import and run WITHOUT a display.

pybfm imports pyglet unconditionally (pybfm/bfm/__init__.py:1 -> bfm.py:4) and issues OpenGL calls even in
headless mode (pybfm/bfm/instance.py:19-83), but none of that is on the compute path.  This package
lets the reference's scripts (lepl1110.py, examples/*.py) run unmodified on a GPU compute box that has
neither pyglet nor a GL context: every gl* function is a no-op, every GL constant is 0, windows do not
open and the event loop returns at once.  It is test tooling, not part of the product.
"""

import sys
import types

options = {}


class _NoOp:
	"""callable that accepts anything and returns 0 - also usable as a context/attribute sink"""

	def __init__(self, name):
		self._name = name

	def __call__(self, *args, **kwargs):
		return 0

	def __bool__(self):
		return True

	def __repr__(self):
		return f"<pyglet stub {self._name}>"


def _make_gl():
	import ctypes

	mod = types.ModuleType("pyglet.gl")

	ctypes_types = {
		"GLuint": ctypes.c_uint, "GLint": ctypes.c_int, "GLfloat": ctypes.c_float, "GLdouble": ctypes.c_double,
		"GLsizei": ctypes.c_int, "GLenum": ctypes.c_uint, "GLchar": ctypes.c_char, "GLboolean": ctypes.c_ubyte,
		"GLubyte": ctypes.c_ubyte, "GLbyte": ctypes.c_byte, "GLushort": ctypes.c_ushort, "GLshort": ctypes.c_short,
		"GLsizeiptr": ctypes.c_ssize_t, "GLintptr": ctypes.c_ssize_t, "GLbitfield": ctypes.c_uint,
	}

	for name, value in ctypes_types.items():
		setattr(mod, name, value)

	class Config:
		def __init__(self, **kwargs):
			self.__dict__.update(kwargs)

	mod.Config = Config

	def __getattr__(name):
		if name.startswith("GL_"):
			return 0

		if name.startswith("gl"):
			return _NoOp(name)

		raise AttributeError(name)

	mod.__getattr__ = __getattr__
	return mod


def _make_window():
	mod = types.ModuleType("pyglet.window")

	class NoSuchConfigException(Exception):
		pass

	class Window:
		def __init__(self, *args, **kwargs):
			self.width = kwargs.get("width", 1280)
			self.height = kwargs.get("height", 720)
			self.config = kwargs.get("config")

		def __getattr__(self, name):  # set_exclusive_mouse, switch_to, flip, clear, close, ...
			if name.startswith("__"):
				raise AttributeError(name)

			return _NoOp(name)

	class _Symbols:
		def __getattr__(self, name):
			return 0 if name.isupper() else _NoOp(name)

	mod.NoSuchConfigException = NoSuchConfigException
	mod.Window = Window
	mod.key = _Symbols()
	mod.mouse = _Symbols()
	return mod


def _make_simple(name, **attrs):
	mod = types.ModuleType(name)

	for key, value in attrs.items():
		setattr(mod, key, value)

	mod.__getattr__ = lambda attr: _NoOp(f"{name}.{attr}")
	return mod


gl = _make_gl()
window = _make_window()
app = _make_simple("pyglet.app", run=_NoOp("app.run"), exit=_NoOp("app.exit"))
clock = _make_simple("pyglet.clock", schedule_interval=_NoOp("clock.schedule_interval"))

for _mod in (gl, window, app, clock):
	sys.modules[_mod.__name__] = _mod
