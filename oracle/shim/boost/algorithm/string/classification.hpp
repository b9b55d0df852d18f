// STUB: included by the reference's engine sources, nothing of it is used on the compiled path
#pragma once
