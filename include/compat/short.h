/* short.h -- drop-in stand-in: included by qkxtm/QKXTM_util.cpp:5, nothing of it is used */
#pragma once
