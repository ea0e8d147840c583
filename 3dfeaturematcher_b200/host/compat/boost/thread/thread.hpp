// compat: <boost/thread/thread.hpp> (tools.h:41 in the reference): nothing of it is used without the GUI thread
